"""Debug / timing aid for the run-shortcut attention kernel at a bench shape: logits of the run kernel against
the per-edge split-fp16 kernel, per-relation error, fallbacks, CUDA-event time of each."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th
from dgl_kgat_b200 import _lib, models
from dgl_kgat_b200.graph import CKGraph
from dgl_kgat_b200.synth import synth_ckg
cfgname = sys.argv[1] if len(sys.argv) > 1 else "amazon-book"
ck = synth_ckg(cfgname); N, R, E = ck.cfg.n_nodes, ck.cfg.n_relations, ck.cfg.n_edges
dev = th.device("cuda", 0)
th.manual_seed(1234)
m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64).to(dev)
g = CKGraph(N, ck.triplets, n_rel=R); dg = g.device_graph(dev); lib = _lib.load()
runs = dg.runs()
n_runs = int(dg.run_arrays["n_runs"].item())
rrp = dg.run_arrays["rel_run_ptr"].cpu().numpy()
print("edges", E, "runs", n_runs, "runs per relation", np.diff(rrp)[:8], "...", np.diff(rrp)[-3:])
def call(kind):
    lg = th.full((E,), float("nan"), device=dev)
    args = (m.entity_embed.weight.data_ptr(), m.relation_embed.weight.data_ptr(), m.W_R.data_ptr(), 64, 64, lg.data_ptr(), _lib.stream_ptr())
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    for rep in range(3):
        e0.record()
        if kind.startswith("runs"):
            vr = th.empty(n_runs, 64, device=dev) if kind == "runs" else None
            _lib.check(lib.kgat_attention_logits_runs(dg.ref, runs, *args[:6], _lib.ptr(vr), args[6]))
        else:
            os.environ["KGAT_ATT_IMPL"] = kind
            _lib.check(lib.kgat_attention_logits(dg.ref, *args))
        e1.record(); th.cuda.synchronize()
    return lg, e0.elapsed_time(e1)
a, ta = call("f16")
print("f16 ms", ta, "fallbacks", lib.kgat_attention_fallbacks(), "err", lib.kgat_umma_error_flag())
c, tc = call("runs_fused")
print("runs_fused ms", tc, "max diff vs f16", float((a - c).abs().max()))
b, tb = call("runs")
print("runs (split) ms", tb, "fallbacks", lib.kgat_attention_fallbacks(), "err", lib.kgat_umma_error_flag())
d = (a - b).abs()
print("nan in runs:", int(th.isnan(b).sum()), "max abs diff", float(d[~th.isnan(d)].max()), "max |logit|", float(a.abs().max()))
rel = dg.arrays["csr_rel"][:E]
bad = (d > 1e-4) | th.isnan(b)
print("bad edges", int(bad.sum()))
if int(bad.sum()):
    cnt = th.bincount(rel[bad].long(), minlength=R).cpu().numpy()
    tot = th.bincount(rel.long(), minlength=R).cpu().numpy()
    print("bad per relation", [(r, int(cnt[r]), int(tot[r])) for r in range(R) if cnt[r]][:50])
