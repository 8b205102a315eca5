#!/usr/bin/env python
"""bench.py -- KGAT attention + aggregation (fwd+bwd) edges/s and eval users/s on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config amazon-book]

One "step"  = one compute_attention(g) (logits +
per-head softmax, forward only - the reference never back-props it) + one gnn(g) forward + one
backward from a BPR loss on B = 10 240 random (u, i+, i-) triples, dropout 0.  edges/s = E / T_step.
Prints ONE JSON line (rank 0).

 value      the step with inputs resident in HBM (batch ids already on the device), CUDA-event
            timed over exactly K steps after W warm-ups, max over ranks.  Single GPU: the step is
            replayed as one CUDA graph (graphed.GraphedCFStep); an eager, instrumented pass of the
            same K steps runs first and provides the per-kernel events, the launch count
            (gpu_launches = kernels per K steps) and eager_ms_per_step.
 e2e        the same step through the public Model API with the batch ids copied from pinned host
            memory and the loss read back every step (the call sequence of kgat.py:153-165).
 roofline   the dominant kernel of the step: algorithmic bytes (SURVEY.md section 8(d) E-model) / its
            mean launch duration from CUDA events recorded around every launch inside the timed
            region, against the measured HBM copy bandwidth of MEASURED_PEAKS.json.
 cpu_baseline  the oracle port (oracle/kgat_ref.py = the reference's models.py op sequence on
            pure-torch message passing) on the host cores, bounded sample.
 --impl reference   times that CPU port alone, same metric/config, and prints the same line (the steps it timed).
Multi-GPU (torchrun): destination-partitioned graph, strong scaling (fixed graph).
Further keys of the line: `roofline.l2_gather` (E-model bytes of the gather kernels against the L2 -> SM
throughput cap - the physical bound while the tables fit L2), `roofline.measured_dram_frac`, `roofline_eval`
(tensor-pipe fraction of the scoring GEMM), `kernels` / `halves` (per-kernel CUDA events of the eager pass),
`skew` (N = 1: the same step on the Zipf-skewed graph of the same size), and at N > 1 `ranks` (rows, edges,
kernel time and barrier wait of every rank), `dist_parity` (the partitioned step against the single-GPU step on
rank 0, outside the timed region), `scale_10x` (BASELINE configs[4]: the 10x graph, step and user-sharded
evaluation), `nvlink` (byte counters when nvidia-smi has them, else the derived bytes of the pushed buffers).
`--skew` runs the whole bench on the skewed graph (diagnostic, not the headline config).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch as th  # noqa: E402

METRIC = "KG attention+aggregation edges/sec (fwd+bwd), amazon-book shape; eval users/s"
BATCH = 10240
REG = 1e-4
K_EVAL = 20


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", default="amazon-book")
    p.add_argument("--cpu-budget-s", type=float, default=25.0, help="CPU-baseline time budget inside the default run")
    p.add_argument("--ref-budget-s", type=float, default=150.0, help="time budget of --impl reference")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-eval", action="store_true")
    p.add_argument("--no-skew", action="store_true", help="skip the Zipf-skewed graph (N = 1)")
    p.add_argument("--skew", action="store_true", help="diagnostic: run the WHOLE bench on the Zipf-skewed graph "
                                                       "(per-kernel table of the skewed step; not the headline config)")
    return p.parse_args()


# ------------------------------------------------------------------------------------------ helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        return False

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def nvlink_tx_rx_kib(gpu_index):
    """Sum over the links of one GPU of the NVLink data counters (`nvidia-smi nvlink -gt d`), in KiB:
    (tx, rx), or None when the tool / counters are not available."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(gpu_index)], capture_output=True,
                             text=True, timeout=20).stdout
    except (OSError, subprocess.TimeoutExpired):
        return None
    tx = rx = 0
    seen = False
    for line in out.splitlines():
        parts = line.replace(":", " ").split()
        if "Tx" in parts and "KiB" in parts:
            tx += int(parts[parts.index("KiB") - 1]); seen = True
        elif "Rx" in parts and "KiB" in parts:
            rx += int(parts[parts.index("KiB") - 1]); seen = True
    return (tx, rx) if seen else None


def measured_tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return float(d["bf16_tflops"]), "measured bf16 burst (MEASURED_PEAKS.json)"
    return 1650.0, "fallback (B200_PROFILING.md)"


TRAFFIC_FILES = ("traffic_r2k.json", "traffic_r2f.json", "traffic_r2.json", "traffic_r1.json")     # newest capture first


def traffic_source(kernel):
    for name in TRAFFIC_FILES:
        try:
            json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"][kernel]
            return "profiles/%s (ncu --set full, same workload)" % name
        except (OSError, KeyError, ValueError):
            continue
    return None


def measured_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed
    `ncu --set full` capture of this workload (profiles/traffic_r*.json, newest first); None if not captured."""
    for name in TRAFFIC_FILES:
        try:
            return float(json.load(open(os.path.join(ROOT, "profiles", name)))["kernels"][kernel]["dram_bytes"])
        except (OSError, KeyError, ValueError):
            continue
    return None


def workload_config(cfg, world=1, work_mb=None, exchange=None):
    c = {"workload": cfg.name + "-shaped synthetic CKG, 1 attention + 1 gnn fwd+bwd per step",
         "nodes": cfg.n_nodes, "edges": cfg.n_edges, "relations": cfg.n_relations, "layers": "64-32-16",
         "batch": BATCH}
    if work_mb is not None:
        c["l2_policy"] = "inputs larger than L2 (%.0f MB touched per step vs 126 MB L2)" % work_mb
        c["parallelism"] = "dst-partition x%d" % world if world > 1 else "single GPU"
        if world > 1 and exchange is not None:
            c["exchange"] = exchange
    return c


def algorithmic_bytes(N, E, R, dims):
    """SURVEY.md section 8(d) E-model (fp32, 4-byte ids, every per-edge gather counted)."""
    d0 = dims[0]
    out = {"attention_logits": E * (4 + 4 + 4 * d0 + 4) + N * 4 * d0 + R * (d0 * d0 + d0) * 4,
           "segment_softmax": E * 12 + N * 4, "edge_weights_to_csc": E * 12}
    for l in range(len(dims) - 1):
        d_in, d_out = dims[l], dims[l + 1]
        out["layer_fwd_%dx%d" % (d_in, d_out)] = E * (8 + 4 * d_in) + N * 4 * d_in
        # backward of layer l: the d_out-wide dh_nb of layer l+1 is gathered by csc_gather (none for the last
        # layer), layer_bwd itself is the dense remainder
        if l + 1 < len(dims) - 1:
            out["csc_gather_%d" % d_out] = E * (8 + 4 * d_out) + N * 4 * 2 * d_out
        out["layer_bwd_%dx%d" % (d_in, d_out)] = N * 4 * (4 * d_in + 3 * d_out)
    out["input_bwd_%d" % d0] = E * (8 + 4 * d0) + N * 4 * 3 * d0
    return out


def step_bytes(N, E, R):
    return E * 1596 + N * 1536 + R * 16640        # SURVEY.md section 8(d) B_step


def make_batches(cfg, n, seed=99):
    rng = np.random.default_rng(seed)
    U, I = cfg.n_users, cfg.n_items
    return [tuple(th.from_numpy(rng.integers(lo, hi, BATCH)).pin_memory() if th.cuda.is_available()
                  else th.from_numpy(rng.integers(lo, hi, BATCH))
                  for lo, hi in ((0, U), (U, U + I), (U, U + I))) for _ in range(n)]


# ------------------------------------------------------------------------------------------ CPU port
def cpu_port_steps(ck, budget_s, steps, warmup=1, seed=1234):
    """Times the oracle port (reference op sequence on pure-torch message passing) on the host cores:
    `warmup` untimed steps, then EXACTLY `steps` timed ones.  Bounded: if (steps + warmup) passes over the
    full graph do not fit the budget, every step works on a sample of COMPLETE destination rows (softmax
    groups stay intact, the entity table stays full size) sized so that they do."""
    from oracle import kgat_ref
    cores = os.cpu_count() or 1
    th.set_num_threads(cores)
    cfg = ck.cfg
    N, R = cfg.n_nodes, cfg.n_relations
    params = kgat_ref.init_params(N, R, seed=seed)
    rng = np.random.default_rng(5)
    batches = make_batches(cfg, 1)
    ids = [b.clone() for b in batches[0]]

    def one_step(trip):
        t0 = time.perf_counter()
        w = kgat_ref.compute_attention(params, trip, N)
        kgat_ref.cf_step(params, trip, w, N, *ids, REG)
        return time.perf_counter() - t0

    # calibrate on 5 % of the rows
    rows = rng.permutation(N)
    frac0 = 0.05
    sel = np.isin(ck.triplets[:, 0], rows[: int(N * frac0)])
    t_cal = one_step(ck.triplets[sel])
    rate = sel.sum() / t_cal
    want_steps = max(1, int(steps))
    frac = min(1.0, budget_s * rate / ((want_steps + warmup) * ck.triplets.shape[0]))
    if frac < 1.0:
        sel = np.isin(ck.triplets[:, 0], rows[: max(1, int(N * frac))])
        trip = ck.triplets[sel]
        sample = "%.1f%% of destination rows (complete rows, %d edges), %d timed steps" % (100 * frac, trip.shape[0], want_steps)
    else:
        trip = ck.triplets
        sample = "full graph (%d edges), %d timed steps" % (trip.shape[0], want_steps)
    for _ in range(warmup):
        one_step(trip)
    times = [one_step(trip) for _ in range(want_steps)]
    t = float(np.mean(times))
    return {"value": trip.shape[0] / t, "unit": "edges/s", "cores": cores, "kind": "port", "sample": sample,
            "ms_per_step": t * 1e3, "edges": int(trip.shape[0]), "steps": want_steps, "warmup": warmup}


def cpu_port_eval(ck, n_users=256):
    """metric.calc_recall_ndcg restated per user (the reference's loop) on a fixed user sample."""
    from oracle import kgat_ref
    from dgl_kgat_b200.synth import pairs_to_user_dict
    cfg = ck.cfg
    th.manual_seed(0)
    emb = th.randn(cfg.n_nodes, 176)
    tr = pairs_to_user_dict(ck.train_pairs, cfg.n_users)
    te_all = pairs_to_user_dict(ck.valid_pairs, cfg.n_users)
    users = list(te_all.keys())[:n_users]
    te = {u: te_all[u] for u in users}
    t0 = time.perf_counter()
    kgat_ref.calc_recall_ndcg(emb, tr, te, ck.item_id_range, K_EVAL)
    return len(users) / (time.perf_counter() - t0)


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port) alone."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dgl_kgat_b200.synth import synth_ckg
    ck = synth_ckg(args.config, seed=1234)
    cfg = ck.cfg
    # exactly --steps timed steps after --warmup untimed ones (the line reports what was run); when that many
    # passes over the full graph would not fit --ref-budget-s, each step is a bounded row sample of it
    res = cpu_port_steps(ck, args.ref_budget_s, args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": "edges/s", "n_gpus": args.gpus,
        "steps": res["steps"], "warmup": res["warmup"], "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(cfg),
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": "edges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference models.py op sequence on pure-torch CPU message passing (DGL itself cannot be installed)",
    }
    emit(line)


def dist_parity_check(model, dk, ck, ids, dev, rank):
    """One partitioned step and, on rank 0, the single-GPU step on the same inputs (reference semantics:
    models.py:146-168, kgat.py:159-161): attention weights must be bit-identical, loss and gradients equal to
    fp32 re-association noise.  Every rank runs the partitioned step; rank 0 alone builds the whole graph."""
    from dgl_kgat_b200.graph import CKGraph
    cfg = ck.cfg
    model.zero_grad(set_to_none=True)
    dk.compute_attention()
    w_dist = dk.attention_weights()
    emb = dk.gnn()
    loss = dk.get_loss(emb, *ids)
    loss.backward()
    th.cuda.synchronize()
    got = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    loss_d = float(loss.item())
    del emb, loss
    model.zero_grad(set_to_none=True)
    if rank != 0:
        return None
    g = CKGraph(cfg.n_nodes, ck.triplets, n_rel=cfg.n_relations)
    g.ndata["id"] = g.ndata["id"].to(dev)
    with th.no_grad():
        w1 = model.compute_attention(g)
    g.edata["w"] = w1
    emb1 = model.gnn(g, g.ndata["id"])
    loss1 = model.get_loss(emb1, *ids)
    loss1.backward()
    th.cuda.synchronize()

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    out = {"att_bit_identical": bool(th.equal(w1, w_dist)), "loss_rel": abs(loss_d - float(loss1.item())) / abs(float(loss1.item())),
           "grad_entity_rel": rel(got["entity_embed.weight"], model.entity_embed.weight.grad)}
    for i, layer in enumerate(model.layers):
        out["grad_W2_%d_rel" % i] = rel(got["layers.%d.res_fc_2.weight" % i], layer.res_fc_2.weight.grad)
    out["grad_W2_rel"] = max(out["grad_W2_%d_rel" % i] for i in range(len(model.layers)))
    del emb1, loss1
    model.zero_grad(set_to_none=True)
    return out


def run_10x(dev, rank, world, barrier, max_over_ranks, no_eval=False, steps=5, warmup=2):
    """BASELINE.json configs[4]: the 10x Amazon-book graph (1.7 M nodes, 36.6 M edges; the 437 MB entity table
    does not fit L2, so the E-model roofline is the physical bound here), destination-partitioned over the
    ranks, a few graph-replayed steps and the user-sharded full-rank top-20 evaluation."""
    from dgl_kgat_b200 import metric, models
    from dgl_kgat_b200.dist import DistKGAT
    from dgl_kgat_b200.graphed import GraphedDistCFStep
    from dgl_kgat_b200.synth import pairs_to_user_dict, synth_ckg
    t_host = time.perf_counter()
    ck = synth_ckg("amazon-book-10x", seed=1234)
    cfg = ck.cfg
    N, E, R = cfg.n_nodes, cfg.n_edges, cfg.n_relations
    th.manual_seed(1234)
    model = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64,
                         reg_lambda_kg=REG, reg_lambda_gnn=REG).to(dev)
    dk = DistKGAT(model, N, ck.triplets, R, dev)
    ids = [b.to(dev) for b in make_batches(cfg, 1)[0]]
    step = GraphedDistCFStep(dk, BATCH)
    for dst, src in zip(step.ids, ids):
        dst.copy_(src)
    for _ in range(warmup):
        step.graph.replay()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step.graph.replay()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    peak, _ = measured_peaks()
    out = {"config": "amazon-book-10x", "nodes": N, "edges": E, "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms, "edges_per_s": E / (ms * 1e-3),
           "step_frac": (step_bytes(N, E, R) / (ms * 1e-3) / 1e9) / (peak * world), "loss": float(step.loss.item())}
    if not no_eval:
        tr = pairs_to_user_dict(ck.train_pairs, cfg.n_users)
        te_all = pairs_to_user_dict(ck.valid_pairs, cfg.n_users)
        users = list(te_all.keys())[rank::world]
        te = {u: te_all[u] for u in users}
        with th.no_grad():
            emb = dk.full_embedding(dk.gnn())
        items = th.as_tensor(ck.item_id_range, device=dev)
        plan = metric.EvalPlan(tr, te, cfg.n_items)
        metric.topk_items(emb, tr, te, items, K_EVAL, plan=plan)
        barrier()
        e0.record()
        metric.topk_items(emb, tr, te, items, K_EVAL, plan=plan)
        e1.record()
        barrier()
        ms_k = max_over_ranks(e0.elapsed_time(e1))
        out["eval_users"] = len(te_all)
        out["eval_kernel_ms"] = ms_k
        out["eval_users_per_s"] = len(te_all) / (ms_k * 1e-3)
        out["eval_tflops_fp32_equivalent"] = 2.0 * len(te_all) * cfg.n_items * emb.shape[1] / (ms_k * 1e-3) / 1e12
    err = max_over_ranks(float(dk.peer.error())) if dk.peer is not None else 0.0
    out["peer_barrier_timeouts"] = int(err)
    out["setup_s"] = time.perf_counter() - t_host
    step = None
    dk.close()
    return out


def run_skew(config, dev, ms_uniform, steps=5, warmup=2):
    """The same step on the Zipf-skewed graph of the same size (synth_ckg(..., skew=True): item popularity and KG
    tails drawn from zipf(1.5), so single rows hold 10^5..10^6 edges, as the hot entities of a real KG do).
    Reported next to the uniform graph's step: the long rows are split across CTAs (include/kgat_b200.h,
    kgat_graph_split_build), so the step should cost about the same per edge."""
    from dgl_kgat_b200 import models
    from dgl_kgat_b200.graph import CKGraph
    from dgl_kgat_b200.graphed import GraphedCFStep
    from dgl_kgat_b200.synth import synth_ckg
    ck = synth_ckg(config, seed=1234, skew=True)
    cfg = ck.cfg
    N, E, R = cfg.n_nodes, cfg.n_edges, cfg.n_relations
    th.manual_seed(1234)
    model = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64,
                         reg_lambda_kg=REG, reg_lambda_gnn=REG).to(dev)
    g = CKGraph(N, ck.triplets, n_rel=R)
    g.ndata["id"] = g.ndata["id"].to(dev)
    dg = g.device_graph(dev)
    ids = [b.to(dev) for b in make_batches(cfg, 1)[0]]
    step = GraphedCFStep(model, g, BATCH, refresh_attention=True)
    for dst, src in zip(step.ids, ids):
        dst.copy_(src)
    for _ in range(warmup):
        step.graph.replay()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    th.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        step.graph.replay()
    e1.record()
    th.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    indeg = np.bincount(ck.triplets[:, 0], minlength=N)
    outdeg = np.bincount(ck.triplets[:, 2], minlength=N)
    return {"config": cfg.name + ", zipf(1.5) tails / items", "edges": E, "steps": steps, "warmup": warmup,
            "ms_per_step": ms, "edges_per_s": E / (ms * 1e-3), "vs_uniform_step": ms / ms_uniform,
            "max_in_degree": int(indeg.max()), "max_out_degree": int(outdeg.max()),
            "rows_split": {"csr": dg.split_counts[1], "csc": dg.split_counts[3]},
            "segments": {"csr": dg.split_counts[0], "csc": dg.split_counts[2]}, "loss": float(step.loss.item())}


# ------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch.distributed as tdist
    from dgl_kgat_b200 import _lib, metric, models
    from dgl_kgat_b200.graph import CKGraph
    from dgl_kgat_b200.synth import pairs_to_user_dict, synth_ckg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not th.cuda.is_available():
        raise SystemExit("bench.py needs a B200 GPU: the KGAT hot path has no CPU fallback")
    th.cuda.set_device(local)
    dev = th.device("cuda", local)
    distributed = world > 1
    if distributed:
        tdist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.load().kgat_device_check(), "kgat_device_check")

    ck = synth_ckg(args.config, seed=1234, skew=args.skew)
    cfg = ck.cfg
    N, E, R = cfg.n_nodes, cfg.n_edges, cfg.n_relations
    th.manual_seed(1234)
    model = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64,
                         reg_lambda_kg=REG, reg_lambda_gnn=REG).to(dev)
    dims = [64, 64, 32, 16]
    n_total = args.steps + args.warmup
    batches = make_batches(cfg, n_total)
    dev_batches = [tuple(b.to(dev) for b in bt) for bt in batches]

    exchange_name = None
    if distributed:
        from dgl_kgat_b200.dist import DistKGAT
        dk = DistKGAT(model, N, ck.triplets, R, dev)
        exchange_name = dk.exchange

        def step(ids):
            model.zero_grad(set_to_none=True)
            dk.compute_attention()
            emb = dk.gnn()
            loss = dk.get_loss(emb, *ids)
            loss.backward()
            return loss
    else:
        g = CKGraph(N, ck.triplets, n_rel=R)
        g.ndata["id"] = g.ndata["id"].to(dev)
        g.device_graph(dev)

        def step(ids):
            model.zero_grad(set_to_none=True)
            with th.no_grad():
                A_w = model.compute_attention(g)
            g.edata["w"] = A_w
            emb = model.gnn(g, g.ndata["id"])
            loss = model.get_loss(emb, *ids)
            loss.backward()
            return loss

    def barrier():
        if distributed:
            tdist.barrier()
        th.cuda.synchronize()

    def max_over_ranks(ms):
        if not distributed:
            return ms
        t = th.tensor([ms], dtype=th.float64, device=dev)
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    rec = _lib.RECORDER
    # ---- instrumented eager pass: per-kernel CUDA events + launch count (same inputs, same process) ----
    for i in range(args.warmup):
        step(dev_batches[i])
    barrier()
    rec.reset(enabled=True, time_kernels=True)
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(dev_batches[args.warmup + i])
    e1.record()
    barrier()
    ms_eager = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = rec.launches
    kernel_ms = {k: float(np.mean(v)) for k, v in rec.durations_ms().items()}
    rec.reset(enabled=False)

    # ---- value: inputs resident in HBM.  Single GPU: the step replayed as one CUDA graph
    # (graphed.GraphedCFStep = the public call sequence, captured), ids already in its device buffers.
    # Multi-GPU, peer exchange (default): the partitioned step - push kernels + device-side barriers, no
    # NCCL call in it - replayed as one CUDA graph per rank (graphed.GraphedDistCFStep).
    # Multi-GPU, KGAT_DIST_EXCHANGE=nccl: the eager DistKGAT calls (KGAT_BENCH_GRAPH_DIST=1 captures the
    # NCCL collectives too; the gain at N=2 was marginal and teardown after capturing NCCL work unreliable).
    graphed = None
    if not distributed:
        from dgl_kgat_b200.graphed import GraphedCFStep
        graphed = GraphedCFStep(model, g, BATCH, refresh_attention=True)
    elif os.environ.get("KGAT_BENCH_GRAPH_DIST", "1" if dk.peer is not None else "0") == "1":
        from dgl_kgat_b200.graphed import GraphedDistCFStep
        graphed = GraphedDistCFStep(dk, BATCH)
    if graphed is not None:
        for dst, src in zip(graphed.ids, dev_batches[0]):
            dst.copy_(src)
        run_value = lambda i: graphed.graph.replay()
    else:
        run_value = lambda i: step(dev_batches[args.warmup + i])
    for i in range(args.warmup):
        run_value(0)
    with ClockSampler(local) as clocks:
        barrier()
        nv0 = nvlink_tx_rx_kib(local) if (distributed and rank == 0) else None
        barrier()
        e0.record()
        for i in range(args.steps):
            run_value(i)
        e1.record()
        barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        ms_step = ms_total / args.steps
        value = E / (ms_step * 1e-3)
        nvlink = None
        if distributed and rank == 0 and nv0 is None:
            # no counters: DERIVED from the buffers the push kernels store (weights of the owned in-edges, owned
            # rows of h_1..h_{L-1}, of every dh_nb and of d_entity; the 3B loss rows and the dW2 slots are small)
            mc = bool(dk.peer is not None and dk.peer.multicast)
            rows0, e0_ = dk.part.r1 - dk.part.r0, dk.dg_in.n_edges
            per_dest = 4 * (e0_ + rows0 * (sum(dims[1:-1]) + sum(dims[:-1]) + dims[0]))
            nvlink = {"unavailable": "nvidia-smi nvlink -gt d reports N/A for every link on this box",
                      "multicast": mc,
                      "derived_tx_MB_per_step_rank0": per_dest * (1 if mc else world - 1) / 1e6,
                      "derived_note": "bytes of rank 0's pushed rows x destinations (one with the NVSwitch "
                                      "multicast, world - 1 with unicast stores)"}
        if nv0 is not None:
            nv1 = nvlink_tx_rx_kib(local)
            if nv1 is not None:
                # measured, not derived: what rank 0's GPU sent / received over NVLink during the K timed steps
                nvlink = {"scope": "rank 0 GPU, all links, the %d timed steps (nvidia-smi nvlink -gt d)" % args.steps,
                          "tx_MB_per_step": (nv1[0] - nv0[0]) * 1024 / 1e6 / args.steps,
                          "rx_MB_per_step": (nv1[1] - nv0[1]) * 1024 / 1e6 / args.steps,
                          "multicast": bool(dk.peer is not None and dk.peer.multicast)}
        barrier()

        # ---- e2e: host batch ids in, loss out, every step (public API) ---------------------------
        if graphed is None:
            def e2e_step(host_ids):
                return step(tuple(b.to(dev, non_blocking=True) for b in host_ids))
        else:
            e2e_step = lambda host_ids: graphed(*host_ids)
        for i in range(min(args.warmup, 3)):
            e2e_step(batches[i]).item()
        barrier()
        e0.record()
        for i in range(args.steps):
            loss_val = e2e_step(batches[args.warmup + i]).item()      # device -> host read of the step's result
        e1.record()
        barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    e2e = {"value": E / (ms_e2e * 1e-3), "unit": "edges/s", "h2d_bytes_per_step": 3 * BATCH * 8,
           "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e, "last_loss": loss_val,
           "api": "dist.DistKGAT (eager)" if graphed is None else
                  ("graphed.GraphedDistCFStep (CUDA graph of the partitioned step, %s exchange)" % dk.exchange
                   if distributed else "graphed.GraphedCFStep (CUDA graph of the public calls)")}

    # ---- roofline of the dominant kernel ------------------------------------------------------
    peak, peak_src = measured_peaks()
    if distributed:
        # rank 0's share: its owned rows, the in-edges (forward / attention) it gathers over
        n_loc, e_loc = dk.part.r1 - dk.part.r0, dk.dg_in.n_edges
        alg = algorithmic_bytes(n_loc, e_loc, R, dims)
    else:
        alg = algorithmic_bytes(N, E, R, dims)
    table = {}
    for name, ms in kernel_ms.items():
        b = alg.get(name)
        table[name] = {"ms": round(ms, 4), "alg_MB": None if b is None else round(b / 1e6, 2),
                       "GBps": None if b is None else round(b / (ms * 1e-3) / 1e9, 1)}
    roofline = None
    modelled = {k: v for k, v in kernel_ms.items() if alg.get(k) is not None}
    if modelled:
        top = max(modelled, key=modelled.get)
        ach = alg[top] / (kernel_ms[top] * 1e-3) / 1e9
        roofline = {"kernel": top, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": None if distributed else measured_traffic(top), "peak_source": peak_src,
                    "launch_ms": kernel_ms[top],
                    "traffic_source": None if distributed else traffic_source(top),
                    "algorithmic_bytes": alg[top],
                    # what the kernel really moved through HBM (ncu dram bytes of the committed capture) over its
                    # live launch time: the E-model counts L2 hits, this one does not
                    "measured_dram_frac": None if (distributed or measured_traffic(top) is None) else
                                          measured_traffic(top) / (kernel_ms[top] * 1e-3) / 1e9 / peak,
                    # whole job: the step's algorithmic bytes over the step time, against world x the peak
                    "step_frac": (step_bytes(N, E, R) / (ms_step * 1e-3) / 1e9) / (peak * world),
                    "step_alg_GB": step_bytes(N, E, R) / 1e9}
        # fractions against the nominal 8 TB/s north_star quotes as well, and the
        # compulsory-traffic C-model (perfect reuse) next to the E-model
        roofline["frac_of_nominal_8TBps"] = ach / 8000.0
        roofline["step_frac_of_nominal_8TBps"] = roofline["step_frac"] * peak / 8000.0
        roofline["step_c_model_GB"] = (E * (12 + 6 * 8) + N * (256 + 2 * 2 * 640)) / 1e9
        # The physical bound of the aggregation kernels at shapes whose tables fit the 126 MB L2 is the L2 -> SM
        # path, not HBM: ~6300 B/clk chip-wide (B300_MICROARCH.md, LTS throughput cap; not re-measured on B200) =
        # 12.4 TB/s at 1965 MHz.  E-model bytes of each gather kernel against that cap:
        l2_cap = 6300.0 * 1965e6 / 1e9
        roofline["l2_gather"] = {
            "cap_GBps": l2_cap, "cap_source": "6300 B/clk LTS cap of the micro-architecture notes x 1965 MHz (not re-measured)",
            "frac": {k: round(alg[k] / (ms * 1e-3) / 1e9 / l2_cap, 3) for k, ms in kernel_ms.items()
                     if alg.get(k) is not None and k.split("_")[0] in ("layer", "csc", "input") and not k.startswith("layer_bwd")}}
        if distributed:
            roofline["scope"] = "rank 0 (%d of %d rows, %d of %d in-edges); step_frac is whole-job" % (n_loc, N, e_loc, E)

    # ---- N > 1: every rank's share and its kernel time (what the partition cost model is calibrated on) ----
    by_rank = None
    if distributed:
        mine = {"rank": rank, "rows": int(dk.part.r1 - dk.part.r0), "in_edges": int(dk.dg_in.n_edges),
                "out_edges": int(dk.dg_out.n_edges),
                "kernel_ms_sum": round(sum(v for k, v in kernel_ms.items() if k != "peer_barrier"), 4),
                "barrier_wait_ms": round(kernel_ms.get("peer_barrier", 0.0) * 7, 4)}
        gathered = [None] * world
        tdist.all_gather_object(gathered, mine)
        by_rank = gathered

    # ---- the two halves of the step and the epoch mix -------------------
    # from the per-call CUDA events of the eager pass: attention = logits + softmax + CSC re-layout, gnn = the rest;
    # one Amazon-book epoch refreshes the attention once and runs 108 CF iterations (kgat.py:139-168)
    halves = None
    if not distributed and kernel_ms:
        att_keys = ("attention_logits", "segment_softmax", "edge_weights_to_csc")
        t_att = sum(kernel_ms.get(k, 0.0) for k in att_keys)
        t_gnn = sum(v for k, v in kernel_ms.items() if k not in att_keys)
        if t_att > 0 and t_gnn > 0:
            halves = {"attention_ms": t_att, "attention_edges_per_s": E / (t_att * 1e-3),
                      "gnn_fwd_bwd_ms": t_gnn, "gnn_edges_per_s": E / (t_gnn * 1e-3),
                      "epoch_mix_edges_per_s": 109 * E / ((t_att + 108 * t_gnn) * 1e-3),
                      "source": "per-call CUDA events of the eager instrumented pass (1 attention : 108 gnn per epoch)"}

    # ---- eval users/s -------------------------------------------------------------------------
    eval_info = None
    roofline_eval = None
    if not args.no_eval:
        tr = pairs_to_user_dict(ck.train_pairs, cfg.n_users)
        te_all = pairs_to_user_dict(ck.valid_pairs, cfg.n_users)
        users = list(te_all.keys())
        if distributed:
            users = users[rank::world]                     # the shard DistKGAT.calc_recall_ndcg gives this rank
        te = {u: te_all[u] for u in users}
        with th.no_grad():
            if distributed:
                emb = dk.full_embedding(dk.gnn())
            else:
                emb = model.gnn(g, g.ndata["id"]).detach()
        items = th.as_tensor(ck.item_id_range, device=dev)
        plan = metric.EvalPlan(tr, te, cfg.n_items)
        metric.topk_items(emb, tr, te, items, K_EVAL, plan=plan)          # warm-up (uploads the plan)
        reps = 5
        barrier()
        e0.record()
        for _ in range(reps):
            topk, _ = metric.topk_items(emb, tr, te, items, K_EVAL, plan=plan)
        e1.record()
        barrier()
        ms_k = max_over_ranks(e0.elapsed_time(e1)) / reps
        # the metric itself through the public call (host dicts in, two floats out); N > 1: the user-sharded
        # path with its 3-scalar all-reduce, so `recall` is the all-ranks value and equals the N = 1 one
        t0 = time.perf_counter()
        if distributed:
            r, n_ = dk.calc_recall_ndcg(emb, tr, te_all, ck.item_id_range, K_EVAL)
        else:
            r, n_ = metric.calc_recall_ndcg(emb, tr, te, items, K_EVAL, plan=plan)
        th.cuda.synchronize()
        wall = time.perf_counter() - t0
        n_users_total = len(te_all)
        flops = 2.0 * n_users_total * cfg.n_items * emb.shape[1]
        eval_info = {"users": n_users_total, "items": cfg.n_items, "dim": int(emb.shape[1]),
                     "users_per_s": n_users_total / (ms_k * 1e-3), "kernel_ms": ms_k,
                     "tflops_fp32_equivalent": flops / (ms_k * 1e-3) / 1e12,
                     "users_per_s_e2e": n_users_total / wall,
                     "recall": r, "ndcg": n_, "recall_scope": "all eval users (all-reduced over ranks)"}
        # tensor-pipe roofline of the scoring GEMM: it is issued as 3 half-precision products per fp32 product
        # (split-fp16, eval_umma.cu), so the tensor pipe sees 3x the fp32-equivalent FLOPs
        tpeak, tsrc = measured_tensor_peak()
        issued = 3.0 * flops / (ms_k * 1e-3) / 1e12
        roofline_eval = {"kernel": "score_topk", "bound": "tensor", "achieved": issued, "peak": tpeak * world,
                         "unit": "TFLOP/s", "frac": issued / (tpeak * world),
                         "achieved_fp32_equivalent": flops / (ms_k * 1e-3) / 1e12, "peak_source": tsrc,
                         "note": "issued = 3 fp16 tensor-core products per fp32-accurate product"}

    # ---- N > 1: the partitioned step against the single-GPU step, outside every timed region --------
    dist_parity = None
    if distributed:
        dist_parity = dist_parity_check(model, dk, ck, dev_batches[0], dev, rank)

    if distributed and dk.peer is not None:
        # a barrier that timed out (a rank fell out of step) invalidates every number above
        peer_err = max_over_ranks(float(dk.peer.error()))
        if peer_err:
            raise SystemExit("peer barrier timed out (waiting for rank %d): no result" % (int(peer_err) - 1))
    value_path = ("eager dist.DistKGAT step" if graphed is None else
                  "CUDA-graph replay of the step, ids resident (eager instrumented pass: eager_ms_per_step)")

    # ---- N > 1: BASELINE.json configs[4], the 10x graph (437 MB table, larger than L2) ---------------
    scale_10x = None
    if distributed and os.environ.get("KGAT_BENCH_10X", "1") == "1" and args.config == "amazon-book":
        graphed = None
        dk.close()
        dk = None
        th.cuda.empty_cache()
        scale_10x = run_10x(dev, rank, world, barrier, max_over_ranks, no_eval=args.no_eval)

    # ---- N = 1: the Zipf-skewed graph of the same size (long rows split across CTAs) -----------------
    skew = None
    if not distributed and not args.no_skew and not args.skew:
        graphed = None
        skew = run_skew(args.config, dev, ms_step)

    # ---- CPU baseline (rank 0, N = 1) ---------------------------------------------------------
    cpu = None
    if rank == 0 and not distributed and not args.no_cpu_baseline:
        res = cpu_port_steps(ck, args.cpu_budget_s, 2, warmup=1)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if not args.no_eval:
            cpu["eval_users_per_s"] = cpu_port_eval(ck, 128)

    if rank == 0:
        work_mb = (N * 4 * (176 + 2 * (64 + 64 + 32) + 64 + 32 + 16) + E * 4 * 12) / 1e6
        line = {
            "metric": METRIC, "value": value, "unit": "edges/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(cfg, world, work_mb, exchange_name),
                           **({"skew": "zipf(1.5) tails / items (diagnostic run, not the headline config)"}
                              if args.skew else {})),
            "e2e": e2e, "gpu_launches": launches, "eager_ms_per_step": ms_eager,
            "value_path": value_path,
            "clocks": clocks.summary(),
            "roofline": roofline, "cpu_baseline": cpu, "kernels": table, "halves": halves,
            "parity_note": "GPU tests hold attention / embedding / loss / gradients to 1e-5 vs the fp64 oracle at this "
                           "shape (tests/test_gpu_parity.py::test_full_size_step_matches_fp64_oracle); weight gradients "
                           "may use 2x the fp32 reference's own error; top-K is bit-exact outside fp32-unresolvable ties",
            # the keys a reader wants first are LAST: the driver keeps the tail of this line
            "ranks": by_rank,
            "eval": eval_info, "roofline_eval": roofline_eval, "dist_parity": dist_parity, "scale_10x": scale_10x, "nvlink": nvlink, "skew": skew,
        }
        emit(line)
    if distributed:
        graphed = None
        if dk is not None:
            dk.close()
        tdist.barrier()
        tdist.destroy_process_group()


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly one JSON line: keep a private handle on it and point fd 1 at stderr, so that
    # anything a native library prints (NCCL's version banner goes to fd 1) cannot end up next to the result
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
