"""ncu target for the evaluation kernels: one wave-filling call of metric.topk_items at Amazon-book shape
(all 53 469 eval users).  Run under `ncu -k regex:score_topk|eval_split|topk_merge ...`; numbers printed under
a profiler are never bench values."""
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
tr = pairs_to_user_dict(ck.train_pairs, cfg.n_users); te = pairs_to_user_dict(ck.valid_pairs, cfg.n_users)
items = th.as_tensor(ck.item_id_range, device=dev)
plan = metric.EvalPlan(tr, te, cfg.n_items)
reps = int(os.environ.get("EVAL_REPS", "2"))
for _ in range(reps):
    metric.topk_items(emb, tr, te, items, 20, plan=plan)
th.cuda.synchronize()
if os.environ.get("EVAL_TIME"):
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        metric.topk_items(emb, tr, te, items, 20, plan=plan)
    e1.record(); th.cuda.synchronize()
    print("topk_items ms:", e0.elapsed_time(e1) / 5)
