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
dk = DistKGAT(m, cfg.n_nodes, ck.triplets, cfg.n_relations, dev)
rng = np.random.default_rng(0); U, I = cfg.n_users, cfg.n_items
ids = [th.as_tensor(rng.integers(lo, hi, 10240), device=dev) for lo, hi in ((0, U), (U, U + I), (U, U + I))]
names = ["attention", "gnn_fwd", "loss", "backward"]
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
if rank == 0:
    print(json.dumps({"world": world, **{n: round(float(np.mean(v)), 4) for n, v in acc.items()},
                      "total": round(float(sum(np.mean(v) for v in acc.values())), 4)}))
tdist.destroy_process_group()
