"""Restatement of the KGAT hot path in plain torch-CPU / numpy.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``) -- this is the checker that
travels to the GPU box, and the "port" timed as ``cpu_baseline`` in
``bench.py``.  Each function follows the reference lines it cites, in the same
operation order, so that in fp32 it reproduces the reference on the mini-DGL
shim to re-association noise and in fp64 it serves as the high-precision band
(SURVEY.md appendix A).  Pinned against the unmodified reference in
``tests/test_oracle.py`` and through ``tests/golden``.

Conventions: ``trip`` is the (E, 3) int (h, r, t) array; edge i runs
src = t[i] (tail) -> dst = h[i] (head), edge id == row (dataset.py:114-119).
``params`` is a dict with the reference's state_dict keys: ``entity_embed.weight``
(N, d), ``relation_embed.weight`` (R, rd), ``W_R`` (R, d, rd),
``layers.{i}.res_fc_2.weight`` (d_out, d_in).
"""
from __future__ import annotations

import math

import numpy as np
import torch as th
import torch.nn.functional as F

LEAKY_SLOPE = 0.01          # F.leaky_relu default, models.py:66
NORM_EPS = 1e-12            # F.normalize default, models.py:165


# --------------------------------------------------------------------------- parameters
def layer_dims(input_dim, n_hidden, n_layers):
    """(d_in, d_out) per KGATConv, models.py:92-106: widths n_hidden // 2**i."""
    dims, d_in = [], input_dim
    for i in range(n_layers):
        d_out = n_hidden // int(math.pow(2, i))
        dims.append((d_in, d_out))
        d_in = d_out
    return dims


def init_params(n_entities, n_relations, input_dim=64, relation_dim=64, n_hidden=64,
                n_layers=3, seed=1234, dtype=th.float32):
    """Same construction ORDER and initialisers as ``Model.__init__`` (models.py:86-106):
    entity Embedding, relation Embedding, W_R xavier_uniform(gain relu), then one
    bias-free Linear per layer.  With ``th.manual_seed(seed)`` this yields the same
    numbers as building the reference Model under that seed."""
    th.manual_seed(seed)
    p = {}
    p["entity_embed.weight"] = th.nn.Embedding(n_entities, input_dim).weight.detach().clone()
    p["relation_embed.weight"] = th.nn.Embedding(n_relations, relation_dim).weight.detach().clone()
    w = th.empty(n_relations, input_dim, relation_dim)
    th.nn.init.xavier_uniform_(w, gain=th.nn.init.calculate_gain("relu"))
    p["W_R"] = w
    for i, (d_in, d_out) in enumerate(layer_dims(input_dim, n_hidden, n_layers)):
        p["layers.%d.res_fc_2.weight" % i] = th.nn.Linear(d_in, d_out, bias=False).weight.detach().clone()
    return {k: v.to(dtype) for k, v in p.items()}


def n_layers_of(params):
    return sum(1 for k in params if k.startswith("layers."))


# --------------------------------------------------------------------------- attention
def att_logits(params, trip):
    """models.py:135-152: for every relation r, over its edges in ascending edge id,
    t_r = e_t W_r, h_r = e_h W_r, logit = t_r . tanh(h_r + e_r)."""
    ent, rel, W_R = params["entity_embed.weight"], params["relation_embed.weight"], params["W_R"]
    h = th.as_tensor(trip[:, 0], dtype=th.long)
    r = th.as_tensor(trip[:, 1], dtype=th.long)
    t = th.as_tensor(trip[:, 2], dtype=th.long)
    out = th.zeros(trip.shape[0], 1, dtype=ent.dtype)
    for i in range(W_R.shape[0]):
        eid = (r == i).nonzero().squeeze(1)                       # filter_edges, models.py:150
        t_r = th.matmul(ent[t[eid]], W_R[i])                      # src = tail, models.py:140
        h_r = th.matmul(ent[h[eid]], W_R[i])                      # dst = head, models.py:141
        a = th.bmm(t_r.unsqueeze(1), th.tanh(h_r + rel[r[eid]]).unsqueeze(2)).squeeze(-1)
        out[eid] = a                                              # apply_edges, models.py:152
    return out


def edge_softmax(logits, dst, n_nodes):
    """Softmax over the in-edges of every destination (head) node, models.py:153."""
    dst = th.as_tensor(dst, dtype=th.long)
    mx = th.full((n_nodes,) + tuple(logits.shape[1:]), -float("inf"), dtype=logits.dtype)
    mx = mx.scatter_reduce(0, dst.view(-1, 1).expand_as(logits), logits, "amax")
    ex = th.exp(logits - mx[dst])
    den = th.zeros_like(mx).index_add_(0, dst, ex)
    return ex / den[dst]


def compute_attention(params, trip, n_nodes):
    """models.py:146-154 -> (E, 1) weights in original edge order."""
    with th.no_grad():
        return edge_softmax(att_logits(params, trip), trip[:, 0], n_nodes)


# --------------------------------------------------------------------------- propagation
def aggregate(h, w, trip, n_nodes):
    """update_all(u_mul_e('h','w'), sum): out[head] = sum_e w[e] * h[tail(e)], models.py:63."""
    src = th.as_tensor(trip[:, 2], dtype=th.long)
    dst = th.as_tensor(trip[:, 0], dtype=th.long)
    return th.zeros(n_nodes, h.shape[1], dtype=h.dtype).index_add(0, dst, h[src] * w)


def kgat_conv(h, w, trip, n_nodes, W2, drop_mask=None, branch=None, info=None):
    """KGATConv.forward, models.py:60-70 (only the h (.) h_N branch exists, :66).

    ``branch`` = (signs, tol), test-only: LeakyReLU's derivative jumps at 0, so a gradient comparison between two
    arithmetics is only meaningful when both took the same branch.  Where a pre-activation lies within the
    forward error band of zero, |pre| <= tol * (|z| |W2|^T) - the rounding of the dot product that produced it,
    relative to the magnitude of its terms - the branch is taken from ``signs`` (the signs of the other
    implementation's layer output) instead of from this evaluation.  ``info`` collects how many elements the two
    implementations decided differently and how many of those lay outside the band."""
    h_n = aggregate(h, w, trip, n_nodes)
    z = th.mul(h, h_n)
    pre = F.linear(z, W2)
    if branch is None:
        out = F.leaky_relu(pre)
    else:
        signs, tol = branch
        with th.no_grad():
            band = tol * F.linear(z.abs(), W2.abs())
            near = pre.abs() <= band
            other_pos = signs > 0
            pos = th.where(near, other_pos, pre > 0)
            if info is not None:
                differ = (other_pos != (pre > 0)) & (signs != 0)
                info.append({"differ": int(differ.sum()), "outside_band": int((differ & ~near).sum()),
                             "in_band": int(near.sum())})
        out = th.where(pos, pre, 0.01 * pre)              # F.leaky_relu's default slope (models.py:66)
    if drop_mask is not None:
        out = out * drop_mask
    return out


def gnn(params, trip, w, n_nodes, drop_masks=None, branch_signs=None, branch_tol=1e-5, info=None):
    """Model.gnn, models.py:156-168: cache [e, norm(h1), norm(h2), ...], un-normalised h
    feeds the next layer, concat along dim 1.  ``branch_signs``: one sign tensor per layer (see kgat_conv)."""
    h = params["entity_embed.weight"]
    cache = [h]
    for i in range(n_layers_of(params)):
        m = None if drop_masks is None else drop_masks[i]
        b = None if branch_signs is None else (branch_signs[i], branch_tol)
        h = kgat_conv(h, w, trip, n_nodes, params["layers.%d.res_fc_2.weight" % i], m, branch=b, info=info)
        cache.append(F.normalize(h, p=2, dim=1))
    return th.cat(cache, 1)


def l2_loss_mean(x):
    return th.mean(th.sum(th.pow(x, 2), dim=1, keepdim=False) / 2.0)       # models.py:9-11


def get_loss(embedding, src_ids, pos_ids, neg_ids, reg_lambda):
    """BPR loss, models.py:170-178."""
    s, p, n = embedding[src_ids], embedding[pos_ids], embedding[neg_ids]
    pos = th.bmm(s.unsqueeze(1), p.unsqueeze(2)).squeeze()
    neg = th.bmm(s.unsqueeze(1), n.unsqueeze(2)).squeeze()
    cf = th.mean(F.logsigmoid(pos - neg)) * (-1.0)
    return cf + reg_lambda * (l2_loss_mean(s) + l2_loss_mean(p) + l2_loss_mean(n))


def transr_loss(params, h, r, pos_t, neg_t, reg_lambda):
    """Model.transR, models.py:114-133 (W_R[r] gathered per sample, :46-47)."""
    ent, rel, W_R = params["entity_embed.weight"], params["relation_embed.weight"], params["W_R"]
    proj = lambda x: F.normalize(th.bmm(x.unsqueeze(1), W_R.index_select(0, r)).squeeze(), p=2, dim=1)
    h_vec, pos_vec, neg_vec = proj(ent[h]), proj(ent[pos_t]), proj(ent[neg_t])
    r_vec = F.normalize(rel[r], p=2, dim=1)
    pos = th.sum(th.pow(h_vec + r_vec - pos_vec, 2), dim=1, keepdim=True)
    neg = th.sum(th.pow(h_vec + r_vec - neg_vec, 2), dim=1, keepdim=True)
    l = th.mean((-1.0) * F.logsigmoid(neg - pos))
    reg = l2_loss_mean(h_vec) + l2_loss_mean(r_vec) + l2_loss_mean(pos_vec) + l2_loss_mean(neg_vec)
    return l + reg_lambda * reg


def layer_signs(embedding, params):
    """Signs of the layer outputs as they appear in an (N, sum d) embedding (normalisation keeps them)."""
    d0 = params["entity_embed.weight"].shape[1]
    out, off = [], d0
    for i in range(n_layers_of(params)):
        d = params["layers.%d.res_fc_2.weight" % i].shape[0]
        out.append(th.sign(embedding[:, off:off + d]).to(params["entity_embed.weight"].dtype))
        off += d
    return out


def cf_step(params, trip, w, n_nodes, src_ids, pos_ids, neg_ids, reg_lambda, branch_signs=None, branch_tol=1e-5,
            info=None):
    """One CF iteration of kgat.py:159-161: gnn forward, BPR loss, backward.
    Returns (loss, embedding, grads) with grads keyed like ``params``.  ``branch_signs``: see kgat_conv."""
    leaves = {k: v.detach().clone().requires_grad_(k == "entity_embed.weight" or k.startswith("layers."))
              for k, v in params.items()}
    emb = gnn(leaves, trip, w, n_nodes, branch_signs=branch_signs, branch_tol=branch_tol, info=info)
    loss = get_loss(emb, src_ids, pos_ids, neg_ids, reg_lambda)
    loss.backward()
    grads = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    return loss.detach(), emb.detach(), grads


# --------------------------------------------------------------------------- evaluation
def one_dcg_at_k(r, k):
    r = np.asarray(r, dtype=np.float64)[:k]                                 # metric.py:15
    return np.sum(r / np.log2(np.arange(2, r.size + 2))) if r.size else 0.0  # method 1, metric.py:20


def one_ndcg_at_k(r, k):
    dcg_max = one_dcg_at_k(sorted(r, reverse=True), k)                      # metric.py:31 (IDCG of the HIT vector)
    return one_dcg_at_k(r, k) / dcg_max if dcg_max else 0.0


def masked_scores(embedding, u_id, train_items, item_ids):
    """metric.py:46-50: scores of one user against every item, train items set to 0.0."""
    score = th.matmul(embedding[u_id], embedding[item_ids].transpose(0, 1))
    score[th.as_tensor(np.asarray(train_items), dtype=th.long)] = 0.0
    return score


def topk_indices(score, k):
    """Descending scores, ties -> lower item index first (the contract; th.sort at
    metric.py:51 is unstable and leaves ties unspecified)."""
    s = score.detach().cpu().numpy()
    return np.argsort(-s, kind="stable")[:k]


def calc_recall_ndcg(embedding, train_user_dict, test_user_dict, item_ids, K):
    """metric.py:36-69, per-user loop.  Returns (recall, ndcg, topk) where topk is the
    (n_eval_users, K) index matrix in dict order."""
    recall_all, ndcg_all, tops = 0.0, 0.0, []
    item_ids = th.as_tensor(np.asarray(item_ids), dtype=th.long)
    with th.no_grad():
        for u_id, pos in test_user_dict.items():
            score = masked_scores(embedding, u_id, train_user_dict[u_id], item_ids)
            rank = topk_indices(score, K)
            tops.append(rank)
            hits = np.zeros(K, dtype=np.float32)
            pos_set = set(int(x) for x in pos)
            for i in range(K):
                if int(rank[i]) in pos_set:
                    hits[i] = 1.0
            if len(pos) > 0:
                recall_all += float(np.sum(hits.astype(np.float64)) / len(pos))     # metric.py:5-7
            ndcg_all += one_ndcg_at_k(hits.tolist(), K)
    n = len(test_user_dict)
    return recall_all / n, ndcg_all / n, np.stack(tops) if tops else np.zeros((0, K), dtype=np.int64)
