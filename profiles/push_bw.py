"""What one rank can push: every rank pushes a contiguous buffer into all arenas at the same time
(kgat_peer_push_buffer: 16-byte stores, unicast to every peer or one multimem.st through the NVSwitch multicast
mapping).  torchrun -n N; KGAT_PEER_MULTICAST=0 selects unicast.  Prints GB/s of payload per rank."""
import os, sys, json
import torch as th, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgl_kgat_b200 import _lib
from dgl_kgat_b200.dist import PeerArena
rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
th.cuda.set_device(local); dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
lib = _lib.load()
MB = 1 << 20
arena = PeerArena(world * 85 * MB + (1 << 16), dev, None)
out = {"world": world, "multicast": arena.multicast, "rows": []}
for mb in (1, 4, 16, 64):
    n = mb * MB // 4
    src = th.randn(n, device=dev)
    dst = arena.buf(("slot", mb), (world, n))[rank]
    st = _lib.stream_ptr()
    for _ in range(3):
        _lib.check(lib.kgat_peer_push_buffer(src.data_ptr(), n, dst.data_ptr(), arena.ref, st), "push")
    arena.barrier(); th.cuda.synchronize(); dist.barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        _lib.check(lib.kgat_peer_push_buffer(src.data_ptr(), n, dst.data_ptr(), arena.ref, st), "push")
    e1.record(); th.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = th.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    arena.barrier(); th.cuda.synchronize()
    ok = bool(th.equal(arena.buf(("slot", mb), (world, n))[rank], src))
    # local copy of the same size for scale
    d2 = th.empty_like(src)
    e0.record()
    for _ in range(reps):
        d2.copy_(src)
    e1.record(); th.cuda.synchronize()
    out["rows"].append({"MB": mb, "push_ms": float(t.item()), "payload_GBps_per_rank": mb * MB / (float(t.item()) * 1e-3) / 1e9,
                        "egress_GBps_per_rank": mb * MB * (1 if arena.multicast else world - 1) / (float(t.item()) * 1e-3) / 1e9,
                        "local_copy_ms": e0.elapsed_time(e1) / reps, "own_copy_ok": ok})
if rank == 0:
    print(json.dumps(out))
arena.close(); dist.destroy_process_group()
