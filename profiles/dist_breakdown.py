"""Per-phase timing of the destination-partitioned step (run under torchrun on N GPUs):
   python -m torch.distributed.run --nproc-per-node N profiles/dist_breakdown.py
Prints, per rank 0, the mean milliseconds of attention / gnn forward / loss / backward."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch as th, torch.distributed as tdist
from dgl_kgat_b200 import models
from dgl_kgat_b200.dist import DistKGAT
from dgl_kgat_b200.synth import synth_ckg

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
th.cuda.set_device(local); dev = th.device("cuda", local)
tdist.init_process_group("nccl", device_id=dev)
ck = synth_ckg("amazon-book"); cfg = ck.cfg
th.manual_seed(1234)
m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=cfg.n_nodes, n_relations=cfg.n_relations, relation_dim=64,
                 reg_lambda_kg=1e-4, reg_lambda_gnn=1e-4).to(dev)
rng = np.random.default_rng(0); U, I = cfg.n_users, cfg.n_items
ids = [th.as_tensor(rng.integers(lo, hi, 10240), device=dev) for lo, hi in ((0, U), (U, U + I), (U, U + I))]
names = ["attention", "gnn_fwd", "loss", "backward"]
for exchange in os.environ.get("KGAT_DIST_EXCHANGES", "nccl,peer").split(","):
    dk = DistKGAT(m, cfg.n_nodes, ck.triplets, cfg.n_relations, dev, exchange=exchange)
    acc = {n: [] for n in names}
    for it in range(13):
        ev = [th.cuda.Event(enable_timing=True) for _ in range(5)]
        m.zero_grad(set_to_none=True)
        tdist.barrier(); th.cuda.synchronize()
        ev[0].record(); dk.compute_attention()
        ev[1].record(); emb = dk.gnn()
        ev[2].record(); loss = dk.get_loss(emb, *ids)
        ev[3].record(); loss.backward()
        ev[4].record(); th.cuda.synchronize()
        if it >= 3:
            for i, n in enumerate(names): acc[n].append(ev[i].elapsed_time(ev[i + 1]))
    del emb, loss
    out = {"world": world, "exchange": exchange, **{n: round(float(np.mean(v)), 4) for n, v in acc.items()},
           "eager_total": round(float(sum(np.mean(v) for v in acc.values())), 4)}
    if dk.peer is not None:
        # the same step as one CUDA graph per rank
        from dgl_kgat_b200.graphed import GraphedDistCFStep
        m.zero_grad(set_to_none=True)
        step = GraphedDistCFStep(dk, 10240)
        for dst, src in zip(step.ids, ids): dst.copy_(src)
        for _ in range(3): step.graph.replay()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        tdist.barrier(); th.cuda.synchronize()
        e0.record()
        for _ in range(20): step.graph.replay()
        e1.record(); th.cuda.synchronize()
        out["graph_replay"] = round(e0.elapsed_time(e1) / 20, 4)
        out["peer_error"] = dk.peer.error()
        step = None
    if rank == 0:
        print(json.dumps(out), flush=True)
    dk.close()
tdist.destroy_process_group()
