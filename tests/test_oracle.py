"""Pins the CPU oracle (oracle/kgat_ref.py, oracle/graph_ref.py):
 (a) against the UNMODIFIED reference models.py / metric.py on the mini-DGL shim (only where
     /root/reference exists: the dev container),
 (b) against the committed golden fixtures generated from that reference,
 (c) against the hand-checked vectors G1-G3 of SURVEY.md section 8(c)."""
import numpy as np
import pytest
import torch as th

from conftest import params_from_golden
from oracle import graph_ref, kgat_ref, ref_loader
from dgl_kgat_b200.synth import synth_ckg

needs_reference = pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present")


# ---------------------------------------------------------------------------------- (c) hand-checked
def test_g1_attention_known_answer(golden_g12):
    p = params_from_golden(golden_g12)
    trip = golden_g12["trip"]
    logits = kgat_ref.att_logits(p, trip).view(-1).numpy()
    # edge 0: (h=0,r=0,t=1): e_t W_0 = [0,1]; tanh(e_h W_0 + e_r) = tanh([1.5,-0.5])
    assert abs(logits[0] - np.tanh(-0.5)) < 1e-7
    np.testing.assert_allclose(logits, [-0.46211717, 0.96402758, -0.46211717, -1.13620687, 0.20243341], atol=1e-7)
    w = kgat_ref.compute_attention(p, trip, 4).view(-1).numpy()
    np.testing.assert_allclose(w, [0.17629719, 0.73385823, 1.0, 0.08984459, 1.0], atol=1e-7)
    np.testing.assert_allclose(w, golden_g12["att_w"].reshape(-1), atol=1e-7)


def test_g2_gnn_known_answer(golden_g12):
    p = params_from_golden(golden_g12)
    trip = golden_g12["trip"]
    w = kgat_ref.compute_attention(p, trip, 4)
    leaves = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    emb = kgat_ref.gnn(leaves, trip, w, 4)
    emb.sum().backward()
    want = [[1, 0, 0.89442712, 0.44721356], [0, 1, -0.00499994, 0.99998748], [1, 1, 0, 0],
            [-1, .5, -0.89442724, -0.44721362]]
    np.testing.assert_allclose(emb.detach().numpy(), want, atol=1e-6)
    assert np.all(emb.detach().numpy()[2, 2:] == 0)          # in-degree-0 node -> zero hidden state
    np.testing.assert_allclose(leaves["entity_embed.weight"].grad.numpy(),
                               [[1.00000024, 2.77365685], [0.98743802, 1], [1.00000012, 1], [0.99999988, 1.00000012]],
                               atol=1e-6)
    np.testing.assert_allclose(emb.detach().numpy(), golden_g12["embedding"], atol=1e-7)


def test_g3_metric_known_answers(golden_g3):
    hv = golden_g3["hit_vectors"]
    want_ndcg = [1.0, 0.6309297535714575, 0.5, 0.9197207891481876, 0.9674679834891693, 0.227670248696953, 0.0]
    for r, want, gold in zip(hv, want_ndcg, golden_g3["ndcg"]):
        got = kgat_ref.one_ndcg_at_k(r.tolist(), 20)
        assert abs(got - want) < 1e-12 and abs(got - gold) < 1e-12
    assert abs(kgat_ref.one_dcg_at_k(hv[4].tolist(), 20) - 2.0616063116448506) < 1e-12
    emb = th.from_numpy(golden_g3["toy_emb"])
    rec, ndcg, top = kgat_ref.calc_recall_ndcg(emb, {0: [0], 1: [4]}, {0: [4, 2], 1: [1, 3]}, np.arange(2, 7), 3)
    assert (rec, ndcg) == (0.75, 1.0)
    assert top.tolist() == [[4, 2, 1], [1, 2, 0]]            # mask-to-0.0 quirk: masked items still rank


# ---------------------------------------------------------------------------------- (b) golden fixtures
def test_oracle_matches_golden_small(golden_small):
    g = golden_small
    p = params_from_golden(g)
    N = int(g["n_nodes"])
    w = kgat_ref.compute_attention(p, g["trip"], N)
    np.testing.assert_allclose(w.numpy(), g["att_w"], rtol=1e-6, atol=1e-8)
    loss, emb, grads = kgat_ref.cf_step(p, g["trip"], th.from_numpy(g["att_w"]), N, th.from_numpy(g["batch_u"]),
                                        th.from_numpy(g["batch_p"]), th.from_numpy(g["batch_n"]), 1e-4)
    np.testing.assert_allclose(emb.numpy(), g["embedding"], rtol=1e-6, atol=1e-7)
    assert abs(loss.item() - float(g["loss"])) < 1e-6
    np.testing.assert_allclose(grads["entity_embed.weight"].numpy(), g["grad_entity"], rtol=1e-5, atol=1e-8)
    for i in range(3):
        np.testing.assert_allclose(grads["layers.%d.res_fc_2.weight" % i].numpy(), g["grad_W2_%d" % i],
                                   rtol=1e-5, atol=1e-8)


def test_oracle_dropout_masks_match_golden(golden_small):
    g = golden_small
    p = params_from_golden(g)
    masks = [th.from_numpy(g["drop_mask_%d" % i]) for i in range(3)]
    emb = kgat_ref.gnn(p, g["trip"], th.from_numpy(g["att_w"]), int(g["n_nodes"]), drop_masks=masks)
    np.testing.assert_allclose(emb.numpy(), g["drop_embedding"], rtol=1e-6, atol=1e-7)


def test_oracle_fp64_band(golden_small):
    """fp32 reference sits within re-association noise of the fp64 restatement."""
    g = golden_small
    p64 = params_from_golden(g, th.float64)
    N = int(g["n_nodes"])
    w64 = kgat_ref.compute_attention(p64, g["trip"], N)
    rel = np.abs(w64.numpy() - g["att_w"]) / np.abs(w64.numpy())
    assert rel.max() < 1e-5
    emb64 = kgat_ref.gnn(p64, g["trip"], w64, N).numpy()
    assert np.abs(emb64 - g["embedding"]).max() / np.abs(emb64).max() < 1e-5


def test_oracle_eval_matches_golden(golden_small, golden_eval):
    e = golden_eval
    users = e["users"]
    tr = {int(u): e["train_items"][e["train_ptr"][j]: e["train_ptr"][j + 1]] for j, u in enumerate(users)}
    te = {int(u): e["test_items"][e["test_ptr"][j]: e["test_ptr"][j + 1]] for j, u in enumerate(users)}
    U, I = int(golden_small["n_users"]), int(golden_small["n_items"])
    emb = th.from_numpy(golden_small["embedding"])
    rec, ndcg, top = kgat_ref.calc_recall_ndcg(emb, tr, te, np.arange(U, U + I), 20)
    assert abs(rec - float(e["recall"])) < 1e-12 and abs(ndcg - float(e["ndcg"])) < 1e-12
    want = np.argsort(-e["masked_scores"], axis=1, kind="stable")[:, :20]
    assert np.array_equal(top, want)


# ---------------------------------------------------------------------------------- (a) the reference itself
@needs_reference
def test_oracle_equals_unmodified_reference_on_tiny():
    from oracle import dgl_shim
    M = ref_loader.models()
    ck = synth_ckg("tiny")
    N, R = ck.cfg.n_nodes, ck.cfg.n_relations
    th.manual_seed(1234)
    model = M.Model(use_KG=True, input_node_dim=64, gnn_model="kgat", num_gnn_layers=3, n_hidden=64, dropout=0.0,
                    n_entities=N, n_relations=R, relation_dim=64, reg_lambda_kg=1e-4, reg_lambda_gnn=1e-4)
    p = kgat_ref.init_params(N, R, seed=1234)
    for k, v in model.state_dict().items():
        assert th.equal(v, p[k]), k
    g = dgl_shim.ShimGraph.from_triplets(N, ck.triplets)
    with th.no_grad():
        w_ref = model.compute_attention(g)
    w = kgat_ref.compute_attention(p, ck.triplets, N)
    assert th.equal(w, w_ref)
    g.edata["w"] = w_ref
    emb_ref = model.gnn(g, g.ndata["id"])
    rng = np.random.default_rng(0)
    u = th.as_tensor(rng.integers(0, ck.cfg.n_users, 512))
    ip = th.as_tensor(rng.integers(ck.cfg.n_users, ck.cfg.n_users + ck.cfg.n_items, 512))
    ineg = th.as_tensor(rng.integers(ck.cfg.n_users, ck.cfg.n_users + ck.cfg.n_items, 512))
    loss_ref = model.get_loss(emb_ref, u, ip, ineg)
    loss_ref.backward()
    loss, emb, grads = kgat_ref.cf_step(p, ck.triplets, w, N, u, ip, ineg, 1e-4)
    assert th.equal(emb, emb_ref.detach()) and loss.item() == loss_ref.item()
    assert th.allclose(grads["entity_embed.weight"], model.entity_embed.weight.grad, rtol=1e-5, atol=1e-8)
    for i in range(3):
        assert th.allclose(grads["layers.%d.res_fc_2.weight" % i], model.layers[i].res_fc_2.weight.grad,
                           rtol=1e-5, atol=1e-8)
    assert model.W_R.grad is None and model.relation_embed.weight.grad is None     # no grad through attention
    # KG-embedding mode
    kb = th.as_tensor(rng.integers(0, N, (4, 256)))
    kr = th.as_tensor(rng.integers(0, R, 256))
    assert th.allclose(model.transR(kb[0], kr, kb[1], kb[2]), kgat_ref.transr_loss(p, kb[0], kr, kb[1], kb[2], 1e-4))


@needs_reference
def test_oracle_metric_equals_unmodified_reference():
    metric = ref_loader.metric()
    rng = np.random.default_rng(5)
    emb = th.from_numpy(rng.standard_normal((300, 16)).astype(np.float32))
    U, I = 100, 200
    tr = {u: np.unique(rng.integers(0, I, rng.integers(1, 30))) for u in range(U)}
    te = {u: np.unique(rng.integers(0, I, rng.integers(1, 6))) for u in range(0, U, 2)}
    want = metric.calc_recall_ndcg(emb, tr, te, th.arange(U, U + I), K=20)
    rec, ndcg, _ = kgat_ref.calc_recall_ndcg(emb, tr, te, np.arange(U, U + I), 20)
    assert abs(rec - want[0]) < 1e-12 and abs(ndcg - want[1]) < 1e-12


@needs_reference
def test_triplet_layout_equals_reference_helpers():
    """dataset.py:165-196 helpers executed from the reference vs the product's restatement."""
    import pandas as pd
    from dgl_kgat_b200 import synth
    from dgl_kgat_b200.dataset import _pairs_to_dict
    D = ref_loader.dataset()
    loader = object.__new__(D.DataLoader)
    rng = np.random.default_rng(2)
    frame = pd.DataFrame({"u": rng.integers(0, 50, 400).astype(np.int32),
                          "v": rng.integers(50, 130, 400).astype(np.int32)}).sort_values(by=["u"], kind="stable")
    want = loader._convert_uv2triplet_np(frame, offset_rel=7, n_uv_rel=1, symmetric=True)
    got = synth.uv_triplets(frame.values, 7)
    assert want.dtype == got.dtype and np.array_equal(want, got)
    d_want = loader._convert_uv_pair2dict(frame)
    d_got = _pairs_to_dict(frame["u"].values, frame["v"].values)
    assert list(d_want.keys()) == list(d_got.keys())
    for k in d_want:
        assert np.array_equal(d_want[k], d_got[k])


# ---------------------------------------------------------------------------------- graph structure oracle
def test_graph_ref_invariants():
    ck = synth_ckg("tiny")
    N, R = ck.cfg.n_nodes, ck.cfg.n_relations
    s = graph_ref.build_structure(ck.triplets, N, R)
    t = ck.triplets
    E = t.shape[0]
    assert s["csr_ptr"][-1] == E and s["csc_ptr"][-1] == E and s["rel_ptr"][-1] == E
    assert np.array_equal(np.sort(s["csr_eid"]), np.arange(E))
    assert np.array_equal(t[s["csr_eid"], 0], s["csr_dst"]) and np.array_equal(t[s["csr_eid"], 2], s["csr_src"])
    assert np.all(np.diff(s["csr_dst"]) >= 0)
    for v in (0, 17, N - 1):                                  # inside a row: (rel, eid) ascending
        a, b = s["csr_ptr"][v], s["csr_ptr"][v + 1]
        key = s["csr_rel"][a:b].astype(np.int64) * E + s["csr_eid"][a:b]
        assert np.all(np.diff(key) > 0)
    assert np.all(np.diff(s["csr_src"][s["csc_pos"]]) >= 0)
    for r in range(R):                                        # bucket r == filter_edges(type == r) as a set
        a, b = s["rel_ptr"][r], s["rel_ptr"][r + 1]
        assert np.array_equal(np.sort(s["csr_eid"][s["bk_pos"][a:b]]), np.flatnonzero(t[:, 1] == r))
    parts = graph_ref.partition_rows(s["csr_ptr"], 4)
    assert parts[0] == 0 and parts[-1] == N and np.all(np.diff(parts) >= 0)
    per = np.diff(s["csr_ptr"][parts])
    assert per.max() - per.min() <= np.diff(s["csr_ptr"]).max() * 2
