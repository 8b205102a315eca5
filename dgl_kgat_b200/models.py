"""KGAT model with the reference's Python surface (reference ``models.py``), backed by the
sm_100a kernels of ``libkgat_b200.so`` instead of DGL message passing.

Kept from the reference, so that ``kgat.py`` and saved state_dicts work unchanged:

* ``Model(use_KG, input_node_dim, gnn_model, num_gnn_layers, n_hidden, dropout, ...)``
  (models.py:73-75) with parameters ``entity_embed.weight``, ``relation_embed.weight``,
  ``W_R`` and ``layers.{i}.res_fc_2.weight`` created in the same order with the same
  initialisers (models.py:86-106) -- the same seed gives the same numbers;
* attention-update mode ``compute_attention(g) -> (E, 1)`` in original edge order
  (models.py:146-154), CF-embedding mode ``gnn(g, x) -> (N, sum d)`` differentiable
  w.r.t. ``entity_embed.weight`` and every ``res_fc_2.weight`` (models.py:156-168),
  ``get_loss`` (models.py:170-178), KG-embedding mode ``transR`` (models.py:114-133);
* ``KGATConv(entity_in_feats, out_feats, dropout, res_type="Bi").forward(g, nfeat)``
  reading ``g.edata['w']`` (models.py:49-70).

Reference behaviour that is reproduced on purpose: only the
``LeakyReLU(W2 (e_h (.) e_N))`` half of the bi-interaction exists (models.py:55,66); no
gradient flows through the attention weights; dropout follows ``self.training``.
Not carried over: the ``graphsage`` branch (needs DGL's SAGEConv) and the dead
``use_KG=False`` branch (models.py:161 references undefined projections).
There is no CPU path: the graph kernels need a B200 and raise otherwise.
"""
from __future__ import annotations

import math

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from . import _lib, ops
from .graph import CKGraph


def _half_sq_norm_mean(*rows):
    """sum over the given (B, d) tensors of mean_b(|row_b|^2 / 2): the L2 term of both losses (models.py:9-11)."""
    return sum(0.5 * v.square().sum(dim=1).mean() for v in rows)


def _bpr_loss_host(embedding, src_ids, pos_ids, neg_ids, reg_lambda):
    """Host statement of models.py:170-178 for CPU tensors (API / oracle tests); the GPU path is ops.BprLoss."""
    s, p, n = (embedding.index_select(0, i.reshape(-1)) for i in (src_ids, pos_ids, neg_ids))
    margin = (s * p).sum(dim=1) - (s * n).sum(dim=1)
    return -F.logsigmoid(margin).mean() + reg_lambda * _half_sq_norm_mean(s, p, n)


def _transr_loss_host(entity, rel_emb, W_R, h, r, pos_t, neg_t, reg_lambda):
    """Host statement of models.py:114-133 for CPU tensors (API / oracle tests); the GPU path is ops.TransRLoss."""
    W = W_R.index_select(0, r)                                               # (B, d, rd), models.py:46-47
    project = lambda ids: F.normalize(th.bmm(entity.index_select(0, ids).unsqueeze(1), W).squeeze(1), p=2, dim=1)
    hv, pv, nv = project(h), project(pos_t), project(neg_t)
    rv = F.normalize(rel_emb.index_select(0, r), p=2, dim=1)
    d_pos = (hv + rv - pv).square().sum(dim=1)
    d_neg = (hv + rv - nv).square().sum(dim=1)
    return -F.logsigmoid(d_neg - d_pos).mean() + reg_lambda * _half_sq_norm_mean(hv, rv, pv, nv)


def bmm_maybe_select(A, B, index):
    """C[i, :] = A[i, :] @ B[index[i]] for dense A, or B[index[i], A[i], :] for integer A
    (same contract as reference models.py:13-47)."""
    if A.dtype == th.int64 and len(A.shape) == 1:
        flat = B.view(-1, B.shape[2])
        return flat.index_select(0, index * B.shape[1] + A)
    return th.bmm(A.unsqueeze(1), B.index_select(0, index)).squeeze()


def _require_graph(g):
    if not isinstance(g, CKGraph):
        raise TypeError("expected a dgl_kgat_b200 CKGraph (see dataset.DataLoader.train_g), got %r" % type(g))


def _dropout_mask(module: nn.Dropout, n, d, device):
    """Scaled keep-mask with the values nn.Dropout would apply to an (n, d) fp32 tensor."""
    if not module.training or module.p == 0.0:
        return None
    return F.dropout(th.ones(n, d, dtype=th.float32, device=device), module.p, True)


class KGATConv(nn.Module):
    def __init__(self, entity_in_feats, out_feats, dropout, res_type="Bi"):
        super(KGATConv, self).__init__()
        self.mess_drop = nn.Dropout(dropout)
        self._res_type = res_type
        if res_type == "Bi":
            self.res_fc_2 = nn.Linear(entity_in_feats, out_feats, bias=False)
        else:
            raise NotImplementedError

    def forward(self, g, nfeat, drop_mask=None):
        """Dropout(LeakyReLU((h (.) h_N) W2^T)), h_N = sum_e w[e] h[src(e)] (models.py:60-70).
        ``drop_mask`` overrides the mask nn.Dropout would draw (used by the parity tests)."""
        _require_graph(g)
        g = g.local_var()
        w_csr, w_csc = g.edge_weight_layouts(nfeat.device)
        dg = g.device_graph(nfeat.device)
        if drop_mask is None:
            drop_mask = _dropout_mask(self.mess_drop, g.n_nodes, self.res_fc_2.out_features, nfeat.device)
        return ops.propagate(ops.PropagatePlan(dg, w_csr, w_csc), [drop_mask], True, nfeat, self.res_fc_2.weight)


class Model(nn.Module):
    def __init__(self, use_KG, input_node_dim, gnn_model, num_gnn_layers, n_hidden, dropout, use_attention=True,
                 n_entities=None, n_relations=None, relation_dim=None,
                 reg_lambda_kg=0.01, reg_lambda_gnn=0.01, res_type="Bi"):
        super(Model, self).__init__()
        if not use_KG:
            raise NotImplementedError("use_KG=False is dead code in the reference (models.py:161)")
        if gnn_model != "kgat":
            raise NotImplementedError("only gnn_model='kgat' is on the B200 hot path "
                                      "(the reference's graphsage branch needs DGL's SAGEConv)")
        self._use_KG = use_KG
        self._n_entities = n_entities
        self._n_relations = n_relations
        self._gnn_model = gnn_model
        self._use_attention = use_attention
        self._reg_lambda_kg = reg_lambda_kg
        self._reg_lambda_gnn = reg_lambda_gnn

        self.entity_embed = nn.Embedding(n_entities, input_node_dim)          # e_h, e_t
        self.relation_embed = nn.Embedding(n_relations, relation_dim)         # e_r
        self.W_R = nn.Parameter(th.Tensor(n_relations, input_node_dim, relation_dim))   # W_r
        nn.init.xavier_uniform_(self.W_R, gain=nn.init.calculate_gain('relu'))

        self.layers = nn.ModuleList()
        for i in range(num_gnn_layers):
            r = int(math.pow(2, i))
            if i == 0:
                self.layers.append(KGATConv(input_node_dim, n_hidden // r, dropout, res_type))
            else:
                r2 = int(math.pow(2, i - 1))
                self.layers.append(KGATConv(n_hidden // r2, n_hidden // r, dropout, res_type))

    # ------------------------------------------------------------------ KG-embedding mode
    def transR(self, h, r, pos_t, neg_t):
        """TransR pairwise ranking loss on L2-normalised projections (models.py:114-133).  On the GPU loss and
        gradients come from the relation-grouped library kernels (``ops.TransRLoss``, 64-wide embeddings; any
        other width is refused - no silent slow path).  CPU tensors get the host statement of the same formula
        (``_transr_loss_host``): the API tests run without a GPU."""
        ent, rel, W = self.entity_embed.weight, self.relation_embed.weight, self.W_R
        if ent.is_cuda:
            if h.dim() != 1 or not ops.transr_supported(ent, rel, W, h.numel()):
                raise _lib.KgatError("transR on the GPU needs fp32 parameters with entity_embed_dim = relation_dim = 64, "
                                     "at most 1024 relations and 1-D index tensors (KGAT_ERR_UNSUPPORTED_DIM); got "
                                     "entity %s, relation %s" % (tuple(ent.shape), tuple(rel.shape)))
            return ops.transr_loss(ent, rel, W, h, r, pos_t, neg_t, self._reg_lambda_kg)
        return _transr_loss_host(ent, rel, W, h, r, pos_t, neg_t, self._reg_lambda_kg)

    # ------------------------------------------------------------------ attention-update mode
    def compute_attention(self, g):
        """pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r) for every edge, softmax over each head's
        in-edges; (E, 1) fp32 in original edge order (models.py:135-154).  No gradient."""
        _require_graph(g)
        dev = self.entity_embed.weight.device
        dg = g.device_graph(dev)
        entity = self._node_features(g)
        with th.no_grad():
            w_eid, w_csr = ops.attention(dg, entity, self.relation_embed.weight, self.W_R)
            w_csc = ops.weights_to_csc(dg, w_eid)
        # `g.edata['w'] = A_w` (kgat.py:144) then finds the CSR/CSC layouts ready
        g._core.remember_weights(w_eid, dev, w_csr, w_csc)
        return w_eid

    # ------------------------------------------------------------------ CF-embedding mode
    def _node_features(self, g):
        if g.ids_are_identity():                       # ndata['id'] = arange(N), dataset.py:118
            return self.entity_embed.weight
        return self.entity_embed(g.ndata['id'].to(self.entity_embed.weight.device))

    def gnn(self, g, x, drop_masks=None):
        """[e, norm(h1), ..., norm(hL)] concatenated, (N, sum d) (models.py:156-168).  ``x`` is
        accepted and ignored like in the reference when use_KG is set (kgat.py:159)."""
        _require_graph(g)
        g = g.local_var()
        dev = self.entity_embed.weight.device
        w_csr, w_csc = g.edge_weight_layouts(dev)
        dg = g.device_graph(dev)
        h = self._node_features(g)
        if drop_masks is None:
            drop_masks = [_dropout_mask(l.mess_drop, g.n_nodes, l.res_fc_2.out_features, dev) for l in self.layers]
        weights = [l.res_fc_2.weight for l in self.layers]
        return ops.propagate(ops.PropagatePlan(dg, w_csr, w_csc), list(drop_masks), False, h, *weights)

    def get_loss(self, embedding, src_ids, pos_dst_ids, neg_dst_ids):
        """BPR loss + L2 regulariser on the gathered rows (models.py:170-178).  On the GPU this is one fused kernel
        whose backward hands ``gnn`` a compact row gradient (widths that are multiples of 4 up to 512; anything
        else is refused).  CPU tensors get the host statement of the formula (``_bpr_loss_host``)."""
        if embedding.is_cuda:
            if not (embedding.dtype == th.float32 and embedding.dim() == 2 and embedding.shape[1] % 4 == 0
                    and embedding.shape[1] <= 512 and src_ids.dim() == 1):
                raise _lib.KgatError("get_loss on the GPU needs an fp32 (N, d) embedding with d a multiple of 4, at most "
                                     "512, and 1-D id tensors (KGAT_ERR_UNSUPPORTED_DIM); got %s %s"
                                     % (tuple(embedding.shape), embedding.dtype))
            return ops.bpr_loss(embedding, src_ids, pos_dst_ids, neg_dst_ids, self._reg_lambda_gnn)
        return _bpr_loss_host(embedding, src_ids, pos_dst_ids, neg_dst_ids, self._reg_lambda_gnn)
