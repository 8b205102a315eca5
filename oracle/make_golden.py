"""Generate tests/golden/*.npz by running the UNMODIFIED reference (models.py, metric.py)
on the mini-DGL shim.  TEST INFRASTRUCTURE; runs only where /root/reference exists.

    python -m oracle.make_golden

Fixtures (inputs AND reference outputs, so the GPU box needs nothing else):
  small_ckg.npz   U=40 users, 120 entities (60 items), 5+2 relations, d=64, layers 64-32-16:
                  attention weights, (N,176) embedding, BPR loss, grads of entity_embed and
                  every res_fc_2 after one CF step; also a dropout run with explicit masks.
  g1_g2.npz       the hand-checkable 4-node vectors G1 (attention) and G2 (1-layer gnn + grad)
                  of SURVEY.md section 8(c), re-derived from the reference code.
  eval_small.npz  metric.calc_recall_ndcg on the small_ckg embedding: recall, ndcg, and the
                  per-user top-20 lists with their masked score vectors.
  g3_metric.npz   metric helper known answers (one_ndcg_at_k etc.) and the 7x2 toy ranking.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch as th

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import dgl_shim, ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def small_ckg(seed=7):
    rng = np.random.default_rng(seed)
    U, ENT, I, RK, T, P = 40, 120, 60, 5, 900, 300
    h = rng.integers(0, ENT, T)
    r = rng.integers(0, RK, T)
    t = rng.integers(0, ENT, T)
    t[:ENT] = np.arange(ENT)
    r[:RK] = np.arange(RK)
    # leave a few entities without in-edges (zero hidden state beyond layer 0)
    h[np.isin(h, [ENT - 1, ENT - 2, ENT - 3])] = 0
    kg = np.stack([h + U, r, t + U], 1).astype(np.int32)
    u = rng.integers(0, U, P)
    v = rng.integers(0, I, P)
    u[:U] = np.arange(U)
    v[:I] = np.arange(I)
    pairs = np.unique(np.stack([u, v + U], 1), axis=0).astype(np.int32)
    n = pairs.shape[0]
    uv = np.zeros((2 * n, 3), np.int32)
    uv[:n, 0], uv[:n, 1], uv[:n, 2] = pairs[:, 0], RK, pairs[:, 1]
    uv[n:, 0], uv[n:, 1], uv[n:, 2] = pairs[:, 1], RK + 1, pairs[:, 0]
    trip = np.vstack([kg, uv])
    # held-out pairs
    hu = rng.integers(0, U, 120)
    hv = rng.integers(0, I, 120)
    held = np.unique(np.stack([hu, hv], 1), axis=0)
    key_t = pairs[:, 0].astype(np.int64) * I + (pairs[:, 1] - U)
    held = held[~np.isin(held[:, 0].astype(np.int64) * I + held[:, 1], key_t)]
    return dict(U=U, ENT=ENT, I=I, R=RK + 2, N=U + ENT, trip=trip, pairs=pairs, held=held.astype(np.int32))


def build_model(M, N, R, dropout=0.0, seed=1234, d=64, hidden=64, layers=3, rd=64):
    th.manual_seed(seed)
    return M.Model(use_KG=True, input_node_dim=d, gnn_model="kgat", num_gnn_layers=layers, n_hidden=hidden,
                   dropout=dropout, n_entities=N, n_relations=R, relation_dim=rd,
                   reg_lambda_kg=1e-4, reg_lambda_gnn=1e-4)


class _FixedMask(th.nn.Module):
    """Stands in for nn.Dropout (models.py:52,70) with a known scaled keep-mask."""

    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask


def to_np(sd):
    return {k: v.detach().numpy().copy() for k, v in sd.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    M, metric = ref_loader.models(), ref_loader.metric()

    # ------------------------------------------------------------------ small_ckg
    ck = small_ckg()
    N, R, U, I = ck["N"], ck["R"], ck["U"], ck["I"]
    model = build_model(M, N, R)
    g = dgl_shim.ShimGraph.from_triplets(N, ck["trip"])
    with th.no_grad():
        w = model.compute_attention(g)
    g.edata["w"] = w
    emb = model.gnn(g, g.ndata["id"])
    rng = np.random.default_rng(3)
    B = 64
    bu = th.as_tensor(rng.integers(0, U, B))
    bp = th.as_tensor(rng.integers(U, U + I, B))
    bn = th.as_tensor(rng.integers(U, U + I, B))
    loss = model.get_loss(emb, bu, bp, bn)
    loss.backward()
    out = {"param." + k: v for k, v in to_np(model.state_dict()).items()}
    out.update(trip=ck["trip"], n_nodes=N, n_rel=R, n_users=U, n_items=I, pairs=ck["pairs"], held=ck["held"],
               att_w=w.numpy(), embedding=emb.detach().numpy(), loss=loss.item(),
               batch_u=bu.numpy(), batch_p=bp.numpy(), batch_n=bn.numpy(),
               grad_entity=model.entity_embed.weight.grad.numpy())
    for i, layer in enumerate(model.layers):
        out["grad_W2_%d" % i] = layer.res_fc_2.weight.grad.numpy()
    assert model.W_R.grad is None and model.relation_embed.weight.grad is None

    # dropout with explicit masks: patch each layer's Dropout to multiply by a fixed mask
    model_d = build_model(M, N, R)
    masks = []
    mrng = th.Generator().manual_seed(11)
    for layer in model_d.layers:
        d_out = layer.res_fc_2.out_features
        m = (th.rand(N, d_out, generator=mrng) >= 0.25).float() / 0.75
        masks.append(m)
        layer.mess_drop = _FixedMask(m)
    emb_d = model_d.gnn(g, g.ndata["id"])
    loss_d = model_d.get_loss(emb_d, bu, bp, bn)
    loss_d.backward()
    out.update(drop_embedding=emb_d.detach().numpy(), drop_loss=loss_d.item(),
               drop_grad_entity=model_d.entity_embed.weight.grad.numpy())
    for i, layer in enumerate(model_d.layers):
        out["drop_mask_%d" % i] = masks[i].numpy()
        out["drop_grad_W2_%d" % i] = layer.res_fc_2.weight.grad.numpy()
    np.savez_compressed(os.path.join(OUT, "small_ckg.npz"), **out)

    # ------------------------------------------------------------------ eval on the small graph
    train_dict, test_dict = {}, {}
    for uu, vv in ck["pairs"]:
        train_dict.setdefault(int(uu), []).append(int(vv) - U)
    for uu, vv in ck["held"]:
        test_dict.setdefault(int(uu), []).append(int(vv))
    train_dict = {k: np.asarray(v) for k, v in train_dict.items()}
    test_dict = {k: np.asarray(v) for k, v in test_dict.items() if k in train_dict}
    item_range = th.arange(U, U + I)
    embd = emb.detach()
    recall, ndcg = metric.calc_recall_ndcg(embd, train_dict, test_dict, item_range, K=20)
    users = np.asarray(list(test_dict.keys()))
    scores = np.zeros((len(users), I), np.float32)
    for j, uu in enumerate(users):
        s = th.matmul(embd[uu], embd[item_range].transpose(0, 1))
        s[train_dict[int(uu)]] = 0.0
        scores[j] = s.numpy()
    np.savez_compressed(
        os.path.join(OUT, "eval_small.npz"), recall=recall, ndcg=ndcg, users=users, masked_scores=scores,
        train_ptr=np.cumsum([0] + [len(train_dict[int(u_)]) for u_ in users]),
        train_items=np.concatenate([train_dict[int(u_)] for u_ in users]),
        test_ptr=np.cumsum([0] + [len(test_dict[int(u_)]) for u_ in users]),
        test_items=np.concatenate([test_dict[int(u_)] for u_ in users]))

    # ------------------------------------------------------------------ G1 / G2
    trip = np.array([(0, 0, 1), (0, 1, 2), (3, 0, 0), (0, 0, 3), (1, 1, 3)], np.int32)
    m1 = build_model(M, 4, 2, d=2, hidden=2, layers=1, rd=2)
    with th.no_grad():
        m1.entity_embed.weight.copy_(th.tensor([[1, 0], [0, 1], [1, 1], [-1, .5]]))
        m1.relation_embed.weight.copy_(th.tensor([[.5, -.5], [0, 1.]]))
        m1.W_R.copy_(th.tensor([[[1, 0], [0, 1.]], [[0, 1], [2, 0.]]]))
        m1.layers[0].res_fc_2.weight.copy_(th.tensor([[1, -1], [.5, 2.]]))
    g1 = dgl_shim.ShimGraph.from_triplets(4, trip)
    g1l = g1.local_var()
    with th.no_grad():
        for i in range(2):
            e = g1l.filter_edges(lambda edges: edges.data['type'] == i)
            m1.W_r = m1.W_R[i]
            g1l.apply_edges(m1._att_score, e)
        logits = g1l.edata["att_w"].clone()
        w1 = m1.compute_attention(g1)
    g1.edata["w"] = w1
    e1 = m1.gnn(g1, g1.ndata["id"])
    e1.sum().backward()
    np.savez_compressed(os.path.join(OUT, "g1_g2.npz"), trip=trip, logits=logits.numpy(), att_w=w1.numpy(),
                        embedding=e1.detach().numpy(), grad_entity=m1.entity_embed.weight.grad.numpy(),
                        grad_W2=m1.layers[0].res_fc_2.weight.grad.numpy(),
                        **{"param." + k: v for k, v in to_np(m1.state_dict()).items()})

    # ------------------------------------------------------------------ G3
    def hv(*ones):
        r = [0.0] * 20
        for i in ones:
            r[i] = 1.0
        return r
    cases = [hv(0), hv(1), hv(2), hv(0, 2), hv(0, 1, 3), hv(19), hv()]
    emb7 = th.tensor([[1, 0], [0, 1], [.9, .1], [.1, .9], [.5, .5], [-1, 0], [.8, -.2]])
    rec7, ndcg7 = metric.calc_recall_ndcg(emb7, {0: [0], 1: [4]}, {0: [4, 2], 1: [1, 3]}, th.arange(2, 7), K=3)
    np.savez_compressed(
        os.path.join(OUT, "g3_metric.npz"), hit_vectors=np.asarray(cases),
        ndcg=np.asarray([metric.one_ndcg_at_k(c, 20) for c in cases]),
        dcg=np.asarray([metric.one_dcg_at_k(c, 20) for c in cases]),
        recall7=np.asarray([metric.one_recall_at_k(c, 20, 7) for c in cases]),
        toy_emb=emb7.numpy(), toy_recall=rec7, toy_ndcg=ndcg7)
    print("wrote", sorted(os.listdir(OUT)))
    print("G1 logits", logits.view(-1).numpy(), "w", w1.view(-1).numpy())
    print("toy", rec7, ndcg7, "small eval", recall, ndcg)


if __name__ == "__main__":
    main()
