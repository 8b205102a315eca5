"""Run the attention kernel a few times at Amazon-book shape (target for `ncu -k regex:att_logits`)."""
import os
import sys

import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dgl_kgat_b200 import models                            # noqa: E402
from dgl_kgat_b200.graph import CKGraph                     # noqa: E402
from dgl_kgat_b200.synth import synth_ckg                   # noqa: E402

dev = th.device("cuda:0")
ck = synth_ckg(sys.argv[1] if len(sys.argv) > 1 else "amazon-book")
N, R, trip = ck.cfg.n_nodes, ck.cfg.n_relations, ck.triplets
m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64).to(dev)
g = CKGraph(N, trip, n_rel=R)
for _ in range(4):
    m.compute_attention(g)
th.cuda.synchronize()
print("done")
