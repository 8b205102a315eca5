"""Evaluation time against the user-shard size on ONE GPU (what a rank of an N-way user-sharded evaluation runs):
metric.topk_items on every N-th eval user of the Amazon-book shape, N = 1, 2, 4, 8.  EVAL_SHARD=N EVAL_REPS=r runs
only that shard r times (ncu target)."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch as th
from dgl_kgat_b200 import metric
from dgl_kgat_b200.synth import synth_ckg, pairs_to_user_dict
ck = synth_ckg("amazon-book"); cfg = ck.cfg
dev = th.device("cuda", 0)
th.manual_seed(0)
emb = th.randn(cfg.n_nodes, 176, device=dev)
tr = pairs_to_user_dict(ck.train_pairs, cfg.n_users); te_all = pairs_to_user_dict(ck.valid_pairs, cfg.n_users)
items = th.as_tensor(ck.item_id_range, device=dev)
only = int(os.environ.get("EVAL_SHARD", "0"))
out = []
for n in ([only] if only else [1, 2, 4, 8]):
    users = list(te_all.keys())[0::n]
    te = {u: te_all[u] for u in users}
    plan = metric.EvalPlan(tr, te, cfg.n_items)
    for _ in range(int(os.environ.get("EVAL_REPS", "3"))):
        metric.topk_items(emb, tr, te, items, 20, plan=plan)
    th.cuda.synchronize()
    if only:
        continue
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        metric.topk_items(emb, tr, te, items, 20, plan=plan)
    e1.record(); th.cuda.synchronize()
    out.append({"shard": n, "users": len(users), "topk_items_ms": e0.elapsed_time(e1) / 10})
if out:
    print(json.dumps(out))
