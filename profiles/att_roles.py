"""Per-role wait times inside the attention kernel (debug build: KGAT_NVCC_DEFINES=-DKGAT_ATT_PROFILE).

    KGAT_NVCC_DEFINES=-DKGAT_ATT_PROFILE python -m dgl_kgat_b200.build --force && python profiles/att_roles.py
"""
import ctypes
import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgl_kgat_b200 import _lib, models                      # noqa: E402
from dgl_kgat_b200.graph import CKGraph                     # noqa: E402
from dgl_kgat_b200.synth import synth_ckg                   # noqa: E402

dev = th.device("cuda:0")
ck = synth_ckg(sys.argv[1] if len(sys.argv) > 1 else "amazon-book")
N, R, trip = ck.cfg.n_nodes, ck.cfg.n_relations, ck.triplets
m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64).to(dev)
g = CKGraph(N, trip, n_rel=R)
for _ in range(3):
    m.compute_attention(g)
th.cuda.synchronize()
lib = ctypes.CDLL(_lib.load()._name)
out = np.zeros((160, 12), dtype=np.uint64)
rc = lib.kgat_debug_att_profile(out.ctypes.data_as(ctypes.c_void_p))
assert rc == 0
out = out[out[:, 7] > 0].astype(np.float64)
tiles = out[:, 7]
names = ["prod total", "prod wait empty", "mma total", "mma wait full", "mma wait acc_empty", "epi total",
         "epi wait acc_full", None, "epi bar.sync", "epi tmem ld wait"]
print("CTAs", len(out), "tiles/CTA", tiles.mean())
for i, n in enumerate(names):
    if n is None:
        continue
    print("%-20s %8.0f clk/tile" % (n, (out[:, i] / tiles).mean()))
