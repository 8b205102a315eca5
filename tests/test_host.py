"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares, the host
mirror of the reference interface behaves like the reference's, and the product refuses to run
without its CUDA path (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch as th

from dgl_kgat_b200 import _lib, dataset, metric, models, synth
from dgl_kgat_b200.graph import CKGraph

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "kgat_b200.h")).read()
    return sorted(set(re.findall(r"KGAT_API\s+[\w\s\*]+?\b(kgat_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 14
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), "libkgat_b200.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms          # the ctypes binding covers the header exactly
    loaded = _lib.load()
    assert loaded.kgat_version() == 1
    assert loaded.kgat_error_string(-2).decode().startswith("unsupported")


def test_ctypes_signatures_match_the_header_prototypes():
    """Every binding in _lib.SIGNATURES has as many arguments as the C prototype, pointers bound as
    pointers and integers / floats with the width the header declares (a mismatch would corrupt the call
    silently)."""
    text = open(os.path.join(ROOT, "include", "kgat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = dict(re.findall(r"KGAT_API\s+[\w\s\*]+?\b(kgat_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S))
    assert sorted(protos) == sorted(_lib.SIGNATURES)
    scalar = {"int": ctypes.c_int, "int32_t": ctypes.c_int32, "int64_t": ctypes.c_int64, "size_t": ctypes.c_size_t,
              "float": ctypes.c_float, "double": ctypes.c_double, "uint64_t": ctypes.c_uint64}
    for name, params in protos.items():
        args = [a.strip() for a in params.split(",")] if params.strip() not in ("", "void") else []
        bound = _lib.SIGNATURES[name][1]
        assert len(args) == len(bound), "%s: header has %d parameters, binding %d" % (name, len(args), len(bound))
        for a, b in zip(args, bound):
            is_ptr = "*" in a or a.startswith("kgat_stream_t")
            if is_ptr:
                assert b is ctypes.c_void_p or issubclass(b, ctypes._Pointer), (name, a, b)
            else:
                ctype = a.replace("const", "").split()[0]
                assert ctypes.sizeof(b) == ctypes.sizeof(scalar[ctype]), (name, a, b)
                assert (b in (ctypes.c_float, ctypes.c_double)) == (ctype in ("float", "double")), (name, a, b)


def test_peer_group_struct_layout_matches_header():
    text = open(os.path.join(ROOT, "include", "kgat_b200.h")).read()
    assert "#define KGAT_MAX_RANKS %d" % _lib.MAX_RANKS in text
    assert "#define KGAT_PEER_HEADER_BYTES %d" % _lib.PEER_HEADER_BYTES in text
    assert "#define KGAT_IPC_HANDLE_BYTES %d" % _lib.IPC_HANDLE_BYTES in text
    body = text[text.index("typedef struct kgat_peer_group {"): text.index("} kgat_peer_group_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    assert [f for f in re.findall(r"(\w+)(?:\[\w+\])?;", body)] == ["rank", "arena_bytes", "arena", "multicast"]
    assert [n for n, _ in _lib.KgatPeerGroup._fields_] == ["world", "rank", "arena_bytes", "arena", "multicast"]
    assert ctypes.sizeof(_lib.KgatPeerGroup) == 4 + 4 + 8 + 8 * _lib.MAX_RANKS + 8


def test_graph_struct_layout_matches_header():
    text = open(os.path.join(ROOT, "include", "kgat_b200.h")).read()
    body = text[text.index("typedef struct kgat_graph {"): text.index("} kgat_graph_t;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\*(\w+);", body)
    assert fields == _lib.GRAPH_ARRAYS + ["split", "split_scratch"]
    names = [n for n, _ in _lib.KgatGraph._fields_]
    assert names == ["n_nodes", "n_edges", "n_rel"] + fields + ["n_csr_seg", "n_csr_long", "n_csc_seg", "n_csc_long"]
    assert re.search(r"int32_t\s+n_csr_seg, n_csr_long, n_csc_seg, n_csc_long;", body)
    assert ctypes.sizeof(_lib.KgatGraph) == 3 * 8 + len(fields) * 8 + 4 * 4
    assert _lib.KgatGraph.split.offset == 3 * 8 + len(_lib.GRAPH_ARRAYS) * 8
    for name, val in (("ROW", _lib.SPLIT_ROW), ("SEG", _lib.SPLIT_SEG), ("SCRATCH_WIDTH", _lib.SPLIT_SCRATCH_WIDTH)):
        assert "#define KGAT_SPLIT_%s %d" % (name, val) in text


def test_split_table_oracle_on_a_power_law_pointer():
    """The segment tables of kgat_graph_split_build restated in numpy (oracle/graph_ref.split_table): rows above
    split_row edges ascending, ceil(deg / split_seg) segments of equal 32-aligned length, in edge order
    (here with thresholds 2048 / 1024; the GPU tests compare the device tables at the header's thresholds)."""
    from oracle import graph_ref
    deg = np.array([0, 5, 2048, 2049, 7, 100000, 1, 4096, 3000], dtype=np.int64)
    ptr = np.concatenate([[0], np.cumsum(deg)])
    t = graph_ref.split_table(ptr, 2048, 1024)
    assert t["long_row"].tolist() == [3, 5, 7, 8]
    assert t["long_nseg"].tolist() == [3, 98, 4, 3]
    assert t["long_seg0"].tolist() == [0, 3, 101, 105]
    for li, row in enumerate(t["long_row"]):
        s0, n = t["long_seg0"][li], t["long_nseg"][li]
        segs = [(t["seg_beg"][s], t["seg_end"][s]) for s in range(s0, s0 + n)]
        assert segs[0][0] == ptr[row] and segs[-1][1] == ptr[row + 1]
        assert all(a[1] == b[0] for a, b in zip(segs, segs[1:]))                # contiguous, in edge order
        assert all(0 < e - b <= 1024 for b, e in segs)
        assert all((b - ptr[row]) % 32 == 0 for b, _ in segs)
        assert (t["seg_row"][s0:s0 + n] == row).all() and (t["seg_long"][s0:s0 + n] == li).all()


def test_no_cpu_fallback():
    dl = dataset.DataLoader.from_synthetic("tiny")
    th.manual_seed(0)
    m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=dl.n_KG_entity, n_relations=dl.n_KG_relation,
                     relation_dim=64)
    with pytest.raises(_lib.KgatError):
        m.compute_attention(dl.train_g)             # parameters on the CPU: refuse, do not emulate
    with pytest.raises(KeyError):
        m.gnn(dl.train_g, None)                     # edata['w'] not set


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dgl_kgat_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            src = open(os.path.join(pkg, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_model_surface_and_state_dict_keys():
    th.manual_seed(1234)
    m = models.Model(True, 64, "kgat", 3, 64, 0.1, n_entities=100, n_relations=5, relation_dim=64,
                     reg_lambda_kg=1e-4, reg_lambda_gnn=1e-4)
    sd = m.state_dict()
    assert list(sd.keys()) == ["W_R", "entity_embed.weight", "relation_embed.weight",
                               "layers.0.res_fc_2.weight", "layers.1.res_fc_2.weight", "layers.2.res_fc_2.weight"]
    assert [tuple(sd["layers.%d.res_fc_2.weight" % i].shape) for i in range(3)] == [(64, 64), (32, 64), (16, 32)]
    assert tuple(sd["W_R"].shape) == (5, 64, 64)
    # same seed -> same numbers as the oracle's restatement of the reference initialisers
    from oracle import kgat_ref
    p = kgat_ref.init_params(100, 5, seed=1234)
    for k in sd:
        assert th.equal(sd[k], p[k]), k
    with pytest.raises(NotImplementedError):
        models.Model(True, 64, "graphsage", 3, 64, 0.1, n_entities=10, n_relations=2, relation_dim=64)
    # the torch-side heads run anywhere
    ids = th.arange(10)
    assert m.transR(ids, ids % 5, ids + 1, ids + 2).ndim == 0
    emb = th.randn(100, 176)
    assert m.get_loss(emb, ids, ids + 10, ids + 20).ndim == 0


def test_get_loss_and_transr_match_oracle():
    from oracle import kgat_ref
    th.manual_seed(3)
    m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=50, n_relations=4, relation_dim=64,
                     reg_lambda_kg=1e-3, reg_lambda_gnn=1e-3)
    p = {k: v.detach() for k, v in m.state_dict().items()}
    emb = th.randn(50, 176)
    a, b, c = th.randint(0, 50, (3, 32))
    assert th.allclose(m.get_loss(emb, a, b, c), kgat_ref.get_loss(emb, a, b, c, 1e-3))
    r = th.randint(0, 4, (32,))
    assert th.allclose(m.transR(a, r, b, c), kgat_ref.transr_loss(p, a, r, b, c, 1e-3))


def test_graph_surface():
    trip = np.array([(0, 0, 1), (0, 1, 2), (3, 0, 0), (0, 0, 3), (1, 1, 3)], np.int32)
    g = CKGraph(4, trip)
    assert g.number_of_nodes() == 4 and g.number_of_edges() == 5 and g.n_rel == 2
    assert th.equal(g.ndata["id"], th.arange(4)) and g.edata["type"].tolist() == [0, 1, 0, 0, 1]
    src, dst = g.all_edges()
    assert src.tolist() == [1, 2, 0, 3, 3] and dst.tolist() == [0, 0, 3, 0, 1]      # src = tail, dst = head
    lv = g.local_var()
    lv.edata["w"] = th.ones(5, 1)
    assert "w" not in g.edata                       # writes on the local view do not leak
    assert g.ids_are_identity()
    with pytest.raises(ValueError):
        CKGraph(3, trip)
    with pytest.raises(ValueError):
        CKGraph(4, trip, n_rel=1)


def test_dataloader_tsv_roundtrip(tmp_path):
    """Files in the reference's on-disk format give the same object as the in-memory path."""
    ck = synth.synth_ckg("tiny")
    U = ck.cfg.n_users
    d = tmp_path / "datasets" / "toy" / "data"
    d.mkdir(parents=True)

    def dump(name, cols, arr):
        with open(d / name, "w") as f:
            f.write("\t".join(cols) + "\n")
            np.savetxt(f, arr, fmt="%d", delimiter="\t")
    tr = ck.train_pairs.copy(); tr[:, 1] -= U
    va = ck.valid_pairs.copy(); va[:, 1] -= U
    kg = ck.kg.copy(); kg[:, 0] -= U; kg[:, 2] -= U
    dump("uv_train.pd", ["u", "v"], tr)
    half = va.shape[0] // 2                     # from_synthetic: first half validation, second half test
    dump("uv_val.pd", ["u", "v"], va[:half])
    dump("uv_test.pd", ["u", "v"], va[half:])
    dump("kg_item.pd", ["h", "r", "t"], kg)
    a = dataset.DataLoader("toy", data_dir=str(d))
    b = dataset.DataLoader.from_synthetic("tiny")
    assert (a.n_users, a.n_items, a.n_KG_entity, a.n_KG_relation) == (b.n_users, b.n_items, b.n_KG_entity, b.n_KG_relation)
    assert (a.n_users, a.n_items, a.n_KG_entity, a.n_KG_relation) == (2000, 2000, 7000, 11)
    assert np.array_equal(a.train_KG_triplet, b.train_KG_triplet) and a.train_KG_triplet.dtype == np.int32
    assert np.array_equal(a.test_KG_triplet, b.test_KG_triplet)
    assert a.n_train_KG_triplet == 80000 and a.n_test_KG_triplet == 80000 + 2 * a.n_valid
    assert np.array_equal(a.item_id_range, np.arange(2000, 4000))
    # triplet layout [KG ; (u, R, v) ; (v, R+1, u)], dataset.py:84-93,165-185
    n = a.n_train
    t = a.train_KG_triplet
    assert np.array_equal(t[:40000], ck.kg)
    assert np.all(t[40000:40000 + n, 1] == 9) and np.all(t[40000 + n:, 1] == 10)
    assert np.array_equal(t[40000:40000 + n, [0, 2]], t[40000 + n:, [2, 0]])
    # dicts hold raw item ids
    u0 = next(iter(a.train_user_dict))
    assert a.train_user_dict[u0].max() < a.n_items
    assert sorted(a.train_valid_user_dict[u0].tolist()) == sorted(
        a.train_user_dict[u0].tolist() + a.valid_user_dict.get(u0, np.zeros(0, int)).tolist())
    assert a.train_g is a.train_g and a.train_g.n_edges == 80000 and a.test_g.n_edges == a.n_test_KG_triplet


def test_samplers_semantics():
    dl = dataset.DataLoader.from_synthetic("tiny")
    t = dl.train_KG_triplet
    edges = set(map(tuple, t[:, [2, 0, 1]].tolist()))
    seen = 0
    for h, r, pos_t, neg_t, fn in dl.KG_sampler(batch_size=30000):
        seen += len(h)
        # "h" is the graph source (= KG tail), "pos_t" the destination (= KG head): SURVEY Q11
        assert all((int(a), int(b), int(c)) in edges for a, b, c in zip(h[:50], pos_t[:50], r[:50]))
        assert neg_t.min() >= 0 and neg_t.max() < dl.n_KG_entity
        chk = [(int(a), int(b), int(c)) in edges for a, b, c in zip(h[:200], neg_t[:200], r[:200])]
        assert chk == [bool(x) for x in fn[:200]]
    assert seen == dl.n_train_KG_triplet                      # one shuffled pass over all edges
    tot = 0
    for users, pos, neg, fn in dl.CF_pair_sampler(batch_size=10240):
        tot += len(users)
        assert len(users) <= 10240 and neg.max() < dl.n_users + dl.n_items
    assert tot == 2 * dl.n_train


@pytest.mark.parametrize("batch", [256, 5000])
def test_unique_mode_samplers_exclusion_rule(batch):
    """pos_mode="unique" branches of both samplers (reference dataset.py:246-266, :294-310): distinct group
    keys per batch while there are enough groups, the positive is a member of its group, the negative never
    is (the reference re-draws until it is not), no false_neg vector, n // batch + 1 batches."""
    dl = dataset.DataLoader.from_synthetic("tiny")
    trip = dl.train_KG_triplet
    by_head = {}
    for h, r, t in trip:
        by_head.setdefault(int(h), set()).add((int(r), int(t)))
    n = 0
    for h, r, pos_t, neg_t, false_neg in dl.KG_sampler(batch, pos_mode="unique", neg_mode="tail"):
        n += 1
        h, r, pos_t, neg_t = (np.asarray(x) for x in (h, r, pos_t, neg_t))
        assert false_neg is None and h.shape == (batch,)
        if batch <= len(by_head):
            assert len(set(h.tolist())) == batch                    # rd.sample: without replacement
        if n <= 3:
            for a, b, c, d in zip(h.tolist(), r.tolist(), pos_t.tolist(), neg_t.tolist()):
                assert (b, c) in by_head[a] and (b, d) not in by_head[a] and 0 <= d < dl.n_KG_entity
    assert n == dl.n_train_KG_triplet // batch + 1
    n = 0
    for users, pos, neg, false_neg in dl.CF_pair_sampler(batch, pos_mode="unique", neg_mode="exclude_pos"):
        n += 1
        users, pos, neg = (np.asarray(x) for x in (users, pos, neg))
        assert false_neg is None
        if batch <= dl.n_users:
            assert len(set(users.tolist())) == batch
        else:
            assert set(users.tolist()) <= set(int(k) for k in dl.train_user_dict)
        for u, p_, q in zip(users.tolist()[:300], pos.tolist()[:300], neg.tolist()[:300]):
            items = set(int(x) for x in dl.train_user_dict[u])
            assert p_ in items and q not in items and 0 <= q < dl.n_items      # RAW item ids, as in the reference
    assert n == dl.n_train // batch + 1
    with pytest.raises(NotImplementedError):
        next(iter(dl.KG_sampler(batch, pos_mode="unique", neg_mode="head")))
    with pytest.raises(NotImplementedError):
        next(iter(dl.CF_pair_sampler(batch, pos_mode="uniform", neg_mode="exclude_pos")))


def test_device_sampler_same_contract_as_host_sampler():
    """dataset.DeviceEdgeBatchSampler (run on the CPU device here) keeps the batch contract."""
    dl = dataset.DataLoader.from_synthetic("tiny")
    dl.use_device_samplers("cpu")
    t = dl.train_KG_triplet
    edges = set(map(tuple, t[:, [2, 0, 1]].tolist()))
    seen = 0
    for h, r, pos_t, neg_t, fn in dl.KG_sampler(batch_size=30000):
        assert all(isinstance(x, th.Tensor) and x.dtype == th.long for x in (h, r, pos_t, neg_t, fn))
        seen += len(h)
        assert all((int(a), int(b), int(c)) in edges for a, b, c in zip(h[:50], pos_t[:50], r[:50]))
        chk = [(int(a), int(b), int(c)) in edges for a, b, c in zip(h[:200], neg_t[:200], r[:200])]
        assert chk == [bool(x) for x in fn[:200]]
        assert int(neg_t.min()) >= 0 and int(neg_t.max()) < dl.n_KG_entity
    assert seen == dl.n_train_KG_triplet
    tot = sum(len(u) for u, p, n, fn in dl.CF_pair_sampler(batch_size=10240))
    assert tot == 2 * dl.n_train
    dl.use_device_samplers(None)
    assert isinstance(next(iter(dl.KG_sampler(2048)))[0], np.ndarray)


def test_metric_helpers_and_plan(golden_g3, golden_eval, golden_small):
    for r, want in zip(golden_g3["hit_vectors"], golden_g3["ndcg"]):
        assert abs(metric.one_ndcg_at_k(r.tolist(), 20) - want) < 1e-12
    for r, want in zip(golden_g3["hit_vectors"], golden_g3["dcg"]):
        assert abs(metric.one_dcg_at_k(r.tolist(), 20) - want) < 1e-12
    for r, want in zip(golden_g3["hit_vectors"], golden_g3["recall7"]):
        assert abs(metric.one_recall_at_k(r.tolist(), 20, 7) - want) < 1e-12
    # metrics_from_topk on the reference's own top-20 lists reproduces its recall / ndcg
    e = golden_eval
    users = e["users"]
    tr = {int(u): e["train_items"][e["train_ptr"][j]: e["train_ptr"][j + 1]] for j, u in enumerate(users)}
    te = {int(u): e["test_items"][e["test_ptr"][j]: e["test_ptr"][j + 1]] for j, u in enumerate(users)}
    plan = metric.EvalPlan(tr, te, int(golden_small["n_items"]))
    top = np.argsort(-e["masked_scores"], axis=1, kind="stable")[:, :20]
    rec, ndcg, _ = metric.metrics_from_topk(top, plan, 20)
    assert abs(rec - float(e["recall"])) < 1e-12 and abs(ndcg - float(e["ndcg"])) < 1e-12
    assert plan.mask_ptr[-1] == plan.mask_items.size and np.all(np.diff(plan.mask_ptr) >= 0)
    # (the on-device reduction, kgat_recall_ndcg, is held against this one in tests/test_gpu_parity.py); its inputs:
    ptr, items, pos_len = plan.positives_csr()
    assert ptr[0] == 0 and ptr[-1] == len(items) and len(ptr) == plan.n_users + 1
    for j, u in enumerate(te):
        assert items[ptr[j]: ptr[j + 1]].tolist() == sorted(set(int(x) for x in te[u]))
        assert pos_len[j] == len(te[u])                     # duplicates counted, as len(test_user_dict[u]) does


def test_synth_shapes():
    for name, (N, E, R) in {"tiny": (7000, 80000, 11)}.items():
        ck = synth.synth_ckg(name)
        assert (ck.cfg.n_nodes, ck.triplets.shape[0], ck.cfg.n_relations) == (N, E, R)
        assert np.unique(ck.triplets[:, [0, 2]]).size == N and np.unique(ck.triplets[:, 1]).size == R
    c = synth.CONFIGS["amazon-book"]
    assert (c.n_nodes, c.n_edges, c.n_relations) == (159251, 3663302, 41)
    c = synth.CONFIGS["last-fm"]
    assert (c.n_nodes, c.n_edges, c.n_relations) == (81832, 4834673, 11)
    c = synth.CONFIGS["yelp2018"]
    assert (c.n_nodes, c.n_edges, c.n_relations) == (136880, 3492324, 44)
    c = synth.CONFIGS["amazon-book-10x"]
    assert (c.n_nodes, c.n_edges, c.n_relations) == (1706790, 36633020, 41)


@pytest.mark.parametrize("d_in,d_out", [(64, 64), (64, 32), (32, 16), (128, 64), (32, 32)])
def test_tensor_core_tile_decomposition_of_the_layer_gemms(d_in, d_out):
    """Index arithmetic of the mma.sync m16n8k8 tile GEMMs in csrc/propagate.cu, restated with the PTX
    fragment layout (lane = 4 g + t: A holds (g, t), (g+8, t), (g, t+4), (g+8, t+4); B holds (k=t, n=g),
    (k=t+4, n=g); C holds (g, 2t), (g, 2t+1), (g+8, 2t), (g+8, 2t+1)): every warp / tile / k-step
    assignment of the forward product z W2^T and of the two backward products dpre^T z and dpre W2 must
    cover each output element exactly once and sum over the right operands."""
    rng = np.random.default_rng(0)
    z, dp = rng.standard_normal((32, d_in)), rng.standard_normal((32, d_out))
    w2 = rng.standard_normal((d_out, d_in))
    lanes = [(l >> 2, l & 3) for l in range(32)]

    def mma_tile(a_elem, b_elem, ksteps):
        acc = np.zeros((16, 8))
        for ks in range(ksteps):
            a, b = np.zeros((16, 8)), np.zeros((8, 8))
            for g, t in lanes:
                a[g, t], a[g + 8, t], a[g, t + 4], a[g + 8, t + 4] = (a_elem(g, 8 * ks + t), a_elem(g + 8, 8 * ks + t),
                                                                      a_elem(g, 8 * ks + t + 4), a_elem(g + 8, 8 * ks + t + 4))
                b[t, g], b[t + 4, g] = b_elem(8 * ks + t, g), b_elem(8 * ks + t + 4, g)
            acc += a @ b
        return acc

    def scatter(out, acc, r0, c0):
        for g, t in lanes:
            for dr, dc, v in ((0, 0, acc[g, 2 * t]), (0, 1, acc[g, 2 * t + 1]), (8, 0, acc[g + 8, 2 * t]),
                              (8, 1, acc[g + 8, 2 * t + 1])):
                assert np.isnan(out[r0 + g + dr, c0 + 2 * t + dc]), "element written twice"
                out[r0 + g + dr, c0 + 2 * t + dc] = v

    # forward: pre (32 x d_out) = z W2^T; pairs p = warp + 8 q -> (mt, nt) = (p / NT, p % NT)
    nt_fwd = d_out // 8
    pre = np.full((32, d_out), np.nan)
    for p in range(2 * nt_fwd):
        mt, nt = divmod(p, nt_fwd)
        acc = mma_tile(lambda m, k: z[16 * mt + m, k], lambda k, n: w2[8 * nt + n, k], d_in // 8)
        scatter(pre, acc, 16 * mt, 8 * nt)
    assert np.allclose(pre, z @ w2.T)
    # backward: dW2 (d_out x d_in) += dpre^T z over the 32 rows; dz (32 x d_in) = dpre W2
    ntn = d_in // 8
    dw = np.full((d_out, d_in), np.nan)
    for p in range((d_out // 16) * ntn):
        mt, nt = divmod(p, ntn)
        acc = mma_tile(lambda m, k: dp[k, 16 * mt + m], lambda k, n: z[k, 8 * nt + n], 4)
        scatter(dw, acc, 16 * mt, 8 * nt)
    assert np.allclose(dw, dp.T @ z)
    dz = np.full((32, d_in), np.nan)
    for p in range(2 * ntn):
        mt, nt = divmod(p, ntn)
        acc = mma_tile(lambda m, k: dp[16 * mt + m, k], lambda k, n: w2[k, 8 * nt + n], d_out // 8)
        scatter(dz, acc, 16 * mt, 8 * nt)
    assert np.allclose(dz, dp @ w2)



def test_attention_plan_rule(monkeypatch):
    """Which logits kernel a graph gets (ops.attention_plan_for): the run shortcut or the per-edge kernel - decided
    from integer counts, so every rank of a partitioned graph (which all-reduces the counts) reaches the decision of
    the single-GPU step; the per-relation marks serve the opt-in hybrid call."""
    from dgl_kgat_b200 import ops
    monkeypatch.delenv("KGAT_ATT_IMPL", raising=False)
    impl, mode = ops.attention_plan_for([100, 90, 0], [10, 80, 0])
    assert impl == "f16" and mode.tolist() == [1, 0, 0]           # per-relation marks; the hybrid kernel is opt-in
    impl, mode = ops.attention_plan_for([100, 90], [10, 30])
    assert impl == "runs" and mode.tolist() == [1, 1]
    impl, mode = ops.attention_plan_for([0, 0], [0, 0])
    assert impl == "runs" and mode.tolist() == [0, 0]
    monkeypatch.setenv("KGAT_ATT_IMPL", "hybrid")
    assert ops.attention_plan_for([100, 90, 0], [10, 80, 0])[0] == "hybrid"
    monkeypatch.setenv("KGAT_ATT_IMPL", "tf32")
    assert ops.attention_plan_for([100, 90, 0], [10, 80, 0])[0] == "tf32"
