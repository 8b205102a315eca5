"""Where one TransR iteration of the KG phase goes (Amazon-book shape, batch 2048)."""
import os
import sys
import time

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgl_kgat_b200 import models                            # noqa: E402
from dgl_kgat_b200.dataset import DataLoader                # noqa: E402
from dgl_kgat_b200.graphed import GraphedKGStep             # noqa: E402

dev = th.device("cuda:0")
dl = DataLoader.from_synthetic("amazon-book")
dl.use_device_samplers(dev)
m = models.Model(True, 64, "kgat", 3, 64, 0.1, n_entities=dl.n_KG_entity, n_relations=dl.n_KG_relation,
                 relation_dim=64).to(dev)
from dgl_kgat_b200.optim import FusedAdam
opt = FusedAdam(m.parameters(), lr=1e-3)
B = 2048


def timed(fn, n=50):
    for _ in range(5):
        fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    th.cuda.synchronize()
    return e0.elapsed_time(e1) / n, (time.perf_counter() - t0) * 1e3 / n


sampler = iter(dl.KG_sampler(batch_size=B))
h, r, p, n_, _ = next(sampler)


def eager():
    m.zero_grad(set_to_none=True)
    m.transR(h, r, p, n_).backward()


print("eager fused transR fwd+bwd   : %.3f ms device, %.3f ms wall" % timed(eager))
step = GraphedKGStep(m, B)
print("graph replay (fwd+bwd)       : %.3f ms device, %.3f ms wall" % timed(lambda: step(h, r, p, n_)))
step(h, r, p, n_)
print("fused Adam step              : %.3f ms device, %.3f ms wall" % timed(opt.step))
print("sampler batch                : %.3f ms device, %.3f ms wall" % timed(lambda: next(sampler)))


def full():
    hh, rr, pp, nn, _ = next(sampler)
    step(hh, rr, pp, nn)
    opt.step()
    opt.zero_grad()


print("full iteration               : %.3f ms device, %.3f ms wall" % timed(full, 200))
