import os, sys
import torch as th
sys.path.insert(0, "/root/repo")
from dgl_kgat_b200 import models
from dgl_kgat_b200.dataset import DataLoader
dev = th.device("cuda:0")
dl = DataLoader.from_synthetic("amazon-book")
dl.use_device_samplers(dev)
m = models.Model(True, 64, "kgat", 3, 64, 0.1, n_entities=dl.n_KG_entity, n_relations=dl.n_KG_relation, relation_dim=64).to(dev)
h, r, p, n_, _ = next(iter(dl.KG_sampler(batch_size=2048)))
for _ in range(3):
    m.zero_grad(set_to_none=True)
    m.transR(h, r, p, n_).backward()
th.cuda.synchronize()
