import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch as th
from dgl_kgat_b200 import metric
from dgl_kgat_b200.synth import synth_ckg, pairs_to_user_dict
ck = synth_ckg("amazon-book"); cfg = ck.cfg
dev = th.device("cuda", 0)
th.manual_seed(0)
emb = th.randn(cfg.n_nodes, 176, device=dev)
tr = pairs_to_user_dict(ck.train_pairs, cfg.n_users); te_all = pairs_to_user_dict(ck.valid_pairs, cfg.n_users)
users = list(te_all.keys())[:148*128]
te = {u: te_all[u] for u in users}
items = th.as_tensor(ck.item_id_range, device=dev)
plan = metric.EvalPlan(tr, te, cfg.n_items)
for _ in range(2):
    metric.topk_items(emb, tr, te, items, 20, plan=plan)
th.cuda.synchronize()
