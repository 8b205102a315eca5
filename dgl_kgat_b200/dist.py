"""Multi-GPU KGAT: 1-D partition by destination entity, one process per GPU.

New capability (the reference is single-GPU, kgat.py:14,69-71).  Every rank holds the replicated
parameters, the CSR of the in-edges of its own contiguous destination range (forward, attention)
and the CSC of the out-edges of the same range (backward).  Ranges are balanced by edge count.

Exchange steps (``torch.distributed``; NCCL over NVLink on the GPUs, gloo in the CPU tests):
  attention   all-reduce(sum) of the zero-filled (E,) weight vector - each edge is written by
              exactly one rank, so the sum is exact;
  forward     after layers 0 .. L-2: all-gather of the owned rows of the (N, d_l) layer output;
  loss        all-reduce of the 3 x B gathered batch rows (each row owned by one rank), loss
              replicated;
  backward    per layer: all-gather of the owned rows of dh_nb, all-reduce of dW2; finally an
              all-gather of the entity-table gradient;
  eval        all-gather of the (N, sum d) embedding, users sharded, 3-scalar all-reduce.
Softmax segments and aggregation rows never straddle ranks, so single- and multi-GPU results agree
to fp32 re-association noise (the per-row summation order is identical; only dW2 is re-associated).
"""
from __future__ import annotations

import numpy as np
import torch as th
import torch.distributed as tdist

from . import metric, ops
from .graph import CKGraph


def partition_rows(indegree: np.ndarray, n_parts: int) -> np.ndarray:
    """Boundaries b[0..n_parts] of contiguous destination ranges with balanced edge counts:
    b[k] is the first row whose prefix edge count reaches k * E / n_parts."""
    ptr = np.zeros(len(indegree) + 1, dtype=np.int64)
    np.cumsum(indegree, out=ptr[1:])
    E = int(ptr[-1])
    targets = (np.arange(n_parts + 1, dtype=np.int64) * E) // n_parts
    b = np.searchsorted(ptr, targets, side="left").astype(np.int64)
    b[0], b[-1] = 0, len(indegree)
    return np.maximum.accumulate(b)


class Partition:
    """The slice of the CKG owned by ``rank``: edge subsets and the two local graphs."""

    def __init__(self, n_nodes, triplets, n_rel, rank, world):
        trip = np.ascontiguousarray(triplets, dtype=np.int32)
        self.n_nodes, self.n_rel, self.n_edges = int(n_nodes), int(n_rel), trip.shape[0]
        self.rank, self.world = rank, world
        self.bounds = partition_rows(np.bincount(trip[:, 0], minlength=n_nodes), world)
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.eid_in = np.flatnonzero((trip[:, 0] >= self.r0) & (trip[:, 0] < self.r1))     # head owned: forward
        self.eid_out = np.flatnonzero((trip[:, 2] >= self.r0) & (trip[:, 2] < self.r1))    # tail owned: backward
        self.g_in = CKGraph(n_nodes, trip[self.eid_in], n_rel=n_rel)
        self.g_out = CKGraph(n_nodes, trip[self.eid_out], n_rel=n_rel)


def exchange_rows(buf: th.Tensor, bounds, group=None):
    """All-gather of row ranges: rank k's rows ``bounds[k]:bounds[k+1]`` of ``buf`` are broadcast to
    everyone, in place.  Ranges are uneven (edge-balanced), hence one broadcast per owner."""
    for k in range(len(bounds) - 1):
        a, b = int(bounds[k]), int(bounds[k + 1])
        if b > a:
            tdist.broadcast(buf[a:b], src=tdist.get_global_rank(group, k) if group is not None else k, group=group)
    return buf


class _BatchRows(th.autograd.Function):
    """rows = embedding[ids] where every row is contributed by its owner and summed over ranks; the
    loss computed from ``rows`` is replicated, so the backward needs no communication: each rank
    scatters the gradient of the rows it owns."""

    @staticmethod
    def forward(ctx, embedding, ids, r0, r1, group):
        own = (ids >= r0) & (ids < r1)
        rows = embedding[ids] * own.unsqueeze(1).to(embedding.dtype)
        tdist.all_reduce(rows, group=group)
        ctx.save_for_backward(ids, own)
        ctx.shape = embedding.shape
        return rows

    @staticmethod
    def backward(ctx, grad_rows):
        ids, own = ctx.saved_tensors
        g = th.zeros(ctx.shape, dtype=grad_rows.dtype, device=grad_rows.device)
        g.index_add_(0, ids[own], grad_rows[own])
        return g, None, None, None, None


class DistKGAT:
    """Drives a replicated :class:`models.Model` over a destination-partitioned graph."""

    def __init__(self, model, n_nodes, triplets, n_rel, device, group=None):
        self.model, self.device, self.group = model, th.device(device), group
        self.rank, self.world = tdist.get_rank(group), tdist.get_world_size(group)
        self.part = Partition(n_nodes, triplets, n_rel, self.rank, self.world)
        self.dg_in = self.part.g_in.device_graph(self.device)
        self.dg_out = self.part.g_out.device_graph(self.device)
        self._eid_in = th.from_numpy(self.part.eid_in).to(self.device)
        self._eid_out = th.from_numpy(self.part.eid_out).to(self.device)
        self.w_full = None
        self._plan = None

    # attention-update mode ---------------------------------------------------------------
    def compute_attention(self):
        """(E, 1) weights in original edge order on every rank (models.py:146-154)."""
        m, p = self.model, self.part
        with th.no_grad():
            w_local, w_csr = ops.attention(self.dg_in, m.entity_embed.weight, m.relation_embed.weight, m.W_R,
                                           rows=(p.r0, p.r1))
            w_full = th.zeros(p.n_edges, 1, dtype=th.float32, device=self.device)
            w_full[self._eid_in] = w_local
            tdist.all_reduce(w_full, group=self.group)
            _, w_csc = self.dg_out.edge_weights(w_full[self._eid_out])
        self.w_full = w_full
        self._plan = ops.PropagatePlan(
            self.dg_in, w_csr, w_csc, dg_bwd=self.dg_out, rows=(p.r0, p.r1),
            exchange=lambda t: exchange_rows(t, p.bounds, self.group),
            reduce=lambda t: tdist.all_reduce(t, group=self.group))
        return w_full

    # CF-embedding mode -------------------------------------------------------------------
    def gnn(self, drop_masks=None):
        """(N, sum d) embedding whose OWNED rows are valid (others zero); differentiable."""
        if self._plan is None:
            raise RuntimeError("call compute_attention() first")
        m = self.model
        weights = [l.res_fc_2.weight for l in m.layers]
        masks = list(drop_masks) if drop_masks is not None else [None] * len(weights)
        return ops.propagate(self._plan, masks, False, m.entity_embed.weight, *weights)

    def get_loss(self, embedding, src_ids, pos_dst_ids, neg_dst_ids):
        """BPR loss of models.py:170-178 on rows gathered from their owners; replicated value."""
        B = src_ids.numel()
        ids = th.cat([src_ids, pos_dst_ids, neg_dst_ids])
        rows = _BatchRows.apply(embedding, ids, self.part.r0, self.part.r1, self.group)
        table = rows            # (3B, d): rows [0,B) sources, [B,2B) positives, [2B,3B) negatives
        ar = th.arange(B, device=ids.device)
        return self.model.get_loss(table, ar, ar + B, ar + 2 * B)

    def full_embedding(self, embedding):
        with th.no_grad():
            return exchange_rows(embedding.detach().clone(), self.part.bounds, self.group)

    # evaluation ----------------------------------------------------------------------------
    def calc_recall_ndcg(self, embedding_full, train_user_dict, test_user_dict, item_id_range, K):
        """metric.calc_recall_ndcg with the eval users sharded over ranks (metric.py:36-69)."""
        users = list(test_user_dict.keys())
        mine = users[self.rank::self.world]
        sub = {u: test_user_dict[u] for u in mine}
        sums = th.zeros(3, dtype=th.float64, device=self.device)
        if mine:
            topk, plan = metric.topk_items(embedding_full, train_user_dict, sub, item_id_range, K,
                                           plan=metric.EvalPlan(train_user_dict, sub, len(item_id_range)))
            rec, ndcg, _ = metric.metrics_from_topk(topk.cpu().numpy(), plan, K)
            sums += th.tensor([rec * len(mine), ndcg * len(mine), len(mine)], dtype=th.float64, device=self.device)
        tdist.all_reduce(sums, group=self.group)
        n = float(sums[2].item())
        return float(sums[0].item()) / n, float(sums[1].item()) / n
