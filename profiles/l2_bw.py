"""L2-resident read bandwidth on this GPU (the bound of the gather kernels once the entity table sits in L2):
repeated reductions / copies of buffers that fit the 126 MB L2, CUDA-event timed."""
import json
import torch as th
dev = th.device("cuda", 0)
out = []
for mb in (16, 32, 40, 64, 96, 256, 1024):
    n = mb * (1 << 20) // 4
    x = th.randn(n, device=dev)
    y = th.empty_like(x)
    for _ in range(5):
        x.sum()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    reps = 50
    e0.record()
    for _ in range(reps):
        x.sum()
    e1.record(); th.cuda.synchronize()
    ms_sum = e0.elapsed_time(e1) / reps
    for _ in range(5):
        y.copy_(x)
    e0.record()
    for _ in range(reps):
        y.copy_(x)
    e1.record(); th.cuda.synchronize()
    ms_copy = e0.elapsed_time(e1) / reps
    out.append({"MB": mb, "sum_read_GBps": mb * (1 << 20) / (ms_sum * 1e-3) / 1e9,
                "copy_read_plus_write_GBps": 2 * mb * (1 << 20) / (ms_copy * 1e-3) / 1e9})
print(json.dumps(out))
