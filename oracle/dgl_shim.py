"""Mini-DGL: the five DGL 0.4.x primitives the reference hot path calls.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  DGL is a third-party
dependency of the reference that is neither vendored nor pinned (API usage
implies 0.4.1 <= dgl < 0.5) and cannot be installed here, so its documented
semantics are restated in plain torch:

* ``add_edges(u, v)``: edge i goes u[i] -> v[i]; the reference passes
  ``u = triplet[:, 2]`` (tail) and ``v = triplet[:, 0]`` (head), so
  src = tail, dst = head and edge id == triplet row (dataset.py:114-119).
* ``filter_edges(pred)``: ascending ids of edges where pred holds (models.py:150).
* ``apply_edges(fn, eids)``: writes the returned columns at those rows, a
  missing column starts as zeros (models.py:152).
* ``edge_softmax(g, logits)``: softmax over the incoming edges of every
  destination node (models.py:153).
* ``update_all(u_mul_e, sum)``: ``out[dst] = sum_e h[src(e)] * w[e]``; nodes
  without in-edges get zeros (models.py:63).
* ``local_var()``: feature writes do not leak to the caller (models.py:61,148,157).
"""
from __future__ import annotations

import sys
import types

import torch as th


def edge_softmax(g, logits):
    dst = g.dst
    mx = th.full((g.N,) + tuple(logits.shape[1:]), -float("inf"), dtype=logits.dtype)
    mx = mx.scatter_reduce(0, dst.view(-1, 1).expand_as(logits), logits, "amax")
    ex = th.exp(logits - mx[dst])
    den = th.zeros_like(mx).index_add_(0, dst, ex)
    return ex / den[dst]


class EdgeBatch:
    def __init__(self, src, dst, data):
        self.src, self.dst, self.data = src, dst, data


class ShimGraph:
    """What ``models.py`` needs from a DGLGraph, nothing more."""

    def __init__(self, n_nodes, src, dst):
        self.N = int(n_nodes)
        self.src = th.as_tensor(src, dtype=th.long)
        self.dst = th.as_tensor(dst, dtype=th.long)
        self.ndata, self.edata = {}, {}

    @classmethod
    def from_triplets(cls, n_nodes, triplets):
        g = cls(n_nodes, triplets[:, 2], triplets[:, 0])
        g.ndata["id"] = th.arange(n_nodes, dtype=th.long)
        g.edata["type"] = th.as_tensor(triplets[:, 1], dtype=th.long)
        return g

    def number_of_edges(self):
        return self.src.numel()

    def local_var(self):
        g = ShimGraph(self.N, self.src, self.dst)
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        return g

    def _batch(self, eid):
        return EdgeBatch({k: v[self.src[eid]] for k, v in self.ndata.items()},
                         {k: v[self.dst[eid]] for k, v in self.ndata.items()},
                         {k: v[eid] for k, v in self.edata.items()})

    def filter_edges(self, pred):
        return pred(self._batch(th.arange(self.src.numel()))).nonzero().squeeze(1)

    def apply_edges(self, func, eid):
        for k, v in func(self._batch(eid)).items():
            if k not in self.edata:
                self.edata[k] = th.zeros((self.src.numel(),) + tuple(v.shape[1:]), dtype=v.dtype)
            self.edata[k][eid] = v

    def update_all(self, mfn, rfn):
        (_, u, e, _), (_, _, out) = mfn, rfn
        msg = self.ndata[u][self.src] * self.edata[e]
        self.ndata[out] = th.zeros(self.N, msg.shape[1], dtype=msg.dtype).index_add(0, self.dst, msg)


def install():
    """Register the stub modules so ``import dgl...`` inside the reference resolves."""
    if "dgl" in sys.modules and getattr(sys.modules["dgl"], "__kgat_shim__", False):
        return
    dgl = types.ModuleType("dgl")
    dgl.__kgat_shim__ = True
    fn = types.ModuleType("dgl.function")
    fn.u_mul_e = lambda u, e, out: ("u_mul_e", u, e, out)
    fn.sum = lambda msg, out: ("sum", msg, out)
    sm = types.ModuleType("dgl.nn.pytorch.softmax")
    sm.edge_softmax = edge_softmax
    cv = types.ModuleType("dgl.nn.pytorch.conv")
    cv.SAGEConv = object
    nn_ = types.ModuleType("dgl.nn")
    nnpt = types.ModuleType("dgl.nn.pytorch")
    dgl.function = fn
    dgl.nn = nn_
    nn_.pytorch = nnpt
    nnpt.softmax = sm
    nnpt.conv = cv
    for name, mod in {"dgl": dgl, "dgl.function": fn, "dgl.nn": nn_, "dgl.nn.pytorch": nnpt,
                      "dgl.nn.pytorch.softmax": sm, "dgl.nn.pytorch.conv": cv}.items():
        sys.modules[name] = mod
