"""Probe: does torch symmetric memory give an NVSwitch multicast mapping on this box?  torchrun -n 2."""
import os, torch as th, torch.distributed as dist
rank = int(os.environ["RANK"]); th.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = th.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm
t = symm.empty(1 << 20, dtype=th.uint8, device=dev)
h = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "buffer_ptrs", [hex(p) for p in h.buffer_ptrs], "multicast_ptr", hex(h.multicast_ptr), "world", h.world_size, flush=True)
dist.barrier(); dist.destroy_process_group()
