"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in dgl_kgat_b200/dist.py: the
destination partition, the uneven row exchange, the loss-row gather with its communication-free
backward, and the distributed propagation ALGORITHM (oracle arithmetic standing in for the CUDA
kernels) against the single-process oracle."""
import os
import socket

import numpy as np
import pytest
import torch as th
import torch.distributed as tdist
import torch.multiprocessing as mp

from oracle import graph_ref, kgat_ref
from dgl_kgat_b200 import dist as kdist
from dgl_kgat_b200.synth import synth_ckg


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def small_graph():
    ck = synth_ckg("tiny")
    rng = np.random.default_rng(0)
    trip = ck.triplets[rng.choice(ck.triplets.shape[0], 6000, replace=False)]
    return trip, ck.cfg.n_nodes, ck.cfg.n_relations


def test_partition_rows_matches_oracle_and_balances():
    trip, N, R = small_graph()
    indeg = np.bincount(trip[:, 0], minlength=N)
    for parts in (1, 2, 4, 8):
        b = kdist.partition_rows(indeg, parts)
        ptr = np.concatenate([[0], np.cumsum(indeg)])
        assert np.array_equal(b, graph_ref.partition_rows(ptr, parts))
        assert b[0] == 0 and b[-1] == N and np.all(np.diff(b) >= 0)
        per = np.diff(ptr[b])
        assert per.sum() == trip.shape[0] and per.max() - per.min() <= 2 * indeg.max()
    # every edge is owned exactly once in each direction
    owned_in, owned_out = [], []
    for r in range(4):
        p = kdist.Partition(N, trip, R, r, 4)
        owned_in.append(p.eid_in); owned_out.append(p.eid_out)
        assert np.all((trip[p.eid_in, 0] >= p.r0) & (trip[p.eid_in, 0] < p.r1))
        assert np.all((trip[p.eid_out, 2] >= p.r0) & (trip[p.eid_out, 2] < p.r1))
        assert p.g_in.n_edges == len(p.eid_in) and p.g_out.n_edges == len(p.eid_out)
    assert np.array_equal(np.sort(np.concatenate(owned_in)), np.arange(trip.shape[0]))
    assert np.array_equal(np.sort(np.concatenate(owned_out)), np.arange(trip.shape[0]))


def test_partition_rows_minmax_is_a_valid_and_better_partition():
    """The cost-model partition (dist.partition_rows_minmax): boundaries cover every row once, are
    monotone, never do worse on their own objective than the edge-balanced split, and a row-dominated
    model degenerates to equal row counts."""
    trip, N, R = small_graph()
    deg = np.bincount(trip[:, 0], minlength=N) + np.bincount(trip[:, 2], minlength=N)
    ptr = np.concatenate([[0], np.cumsum(deg)])

    def objective(b, row_ns, edge_ns):
        return row_ns * np.diff(b).max() + edge_ns * np.diff(ptr[b]).max()

    for parts in (1, 2, 3, 4, 8):
        for row_ns, edge_ns in ((1.0, 0.18), (9.7, 0.18), (0.01, 1.0)):
            b = kdist.partition_rows_minmax(deg, parts, row_ns, edge_ns)
            assert len(b) == parts + 1 and b[0] == 0 and b[-1] == N and np.all(np.diff(b) >= 0)
            base = kdist.partition_rows(deg, parts)
            assert objective(b, row_ns, edge_ns) <= objective(base, row_ns, edge_ns) * (1 + 1e-9)
        b = kdist.partition_rows_minmax(deg, parts, 1e6, 1e-6)
        assert np.diff(b).max() == -(-N // parts)
    # deterministic (every rank must compute the same boundaries)
    assert np.array_equal(kdist.partition_rows_minmax(deg, 4, 9.7, 0.18), kdist.partition_rows_minmax(deg, 4, 9.7, 0.18))
    # the Partition object uses it when a row cost is given and still owns every edge exactly once
    owned = []
    for r in range(4):
        p = kdist.Partition(N, trip, R, r, 4, row_ns=9.7)
        owned.append(p.eid_in)
        assert np.array_equal(p.bounds, kdist.partition_rows_minmax(deg, 4, 9.7, 0.18))
    assert np.array_equal(np.sort(np.concatenate(owned)), np.arange(trip.shape[0]))


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        th.manual_seed(0)
        trip, N, R = small_graph()
        part = kdist.Partition(N, trip, R, rank, world)
        # --- exchange_rows: uneven ranges, in place
        buf = th.zeros(N, 3)
        buf[part.r0:part.r1] = th.arange(part.r0, part.r1, dtype=th.float32).unsqueeze(1) + th.tensor([0., .5, .25])
        kdist.exchange_rows(buf, part.bounds)
        want = th.arange(N, dtype=th.float32).unsqueeze(1) + th.tensor([0., .5, .25])
        assert th.equal(buf, want)

        # --- distributed attention + propagation with oracle arithmetic on the local sub-graphs
        p = kgat_ref.init_params(N, R, seed=3, dtype=th.float64)
        sub_in = trip[part.eid_in]
        w_local = kgat_ref.compute_attention(p, sub_in, N)              # complete rows -> same softmax groups
        w_full = th.zeros(trip.shape[0], 1, dtype=th.float64)
        w_full[part.eid_in] = w_local
        tdist.all_reduce(w_full)
        w_single = kgat_ref.compute_attention(p, trip, N)
        assert th.allclose(w_full, w_single, rtol=0, atol=1e-15)

        h = p["entity_embed.weight"]
        cache = [h]
        own = slice(part.r0, part.r1)
        for l in range(3):
            out_local = kgat_ref.kgat_conv(h, w_local, sub_in, N, p["layers.%d.res_fc_2.weight" % l])
            nxt = th.zeros_like(out_local)
            nxt[own] = out_local[own]                                   # rows of other ranks: not ours to write
            if l < 2:
                kdist.exchange_rows(nxt, part.bounds)
            cache.append(th.nn.functional.normalize(nxt, p=2, dim=1))
            h = nxt
        emb_local = th.cat(cache, 1)
        emb_single = kgat_ref.gnn(p, trip, w_single, N)
        assert th.allclose(emb_local[own], emb_single[own], rtol=1e-12, atol=1e-14)

        # --- attention plan: (head, relation) runs never straddle ranks (a rank owns whole head rows), so the
        # per-relation edge / run counts of the local in-edge graphs add up to the global ones and every rank
        # reaches the single-GPU decision (dist.DistKGAT all-reduces exactly these counts)
        from oracle import graph_ref
        from dgl_kgat_b200 import ops

        def rel_counts(t):
            st = graph_ref.build_structure(t, N, R)
            runs = graph_ref.build_runs(st, R)
            return np.diff(st["rel_ptr"].astype(np.int64)), np.diff(runs["rel_run_ptr"].astype(np.int64))
        e_loc, r_loc = rel_counts(sub_in)
        counts = th.from_numpy(np.stack([e_loc, r_loc]))
        tdist.all_reduce(counts)
        e_glob, r_glob = rel_counts(trip)
        assert np.array_equal(counts[0].numpy(), e_glob) and np.array_equal(counts[1].numpy(), r_glob)
        impl, mode = ops.attention_plan_for(counts[0].numpy(), counts[1].numpy())
        impl1, mode1 = ops.attention_plan_for(e_glob, r_glob)
        assert impl == impl1 and np.array_equal(mode, mode1)

        # --- loss rows: owner contributes, all-reduce, backward scatters locally
        emb = emb_single.clone()
        mask = th.zeros(N, 1, dtype=th.float64); mask[own] = 1
        emb_owned = (emb * mask).requires_grad_(True)
        ids = th.randint(0, N, (96,), generator=th.Generator().manual_seed(5))
        rows = kdist._BatchRows.apply(emb_owned, ids, part.r0, part.r1, None)
        assert th.allclose(rows, emb[ids])
        gsel = th.randn(96, emb.shape[1], dtype=th.float64, generator=th.Generator().manual_seed(6))
        rows.backward(gsel)
        want_g = th.zeros_like(emb).index_add_(0, ids, gsel) * mask
        assert th.allclose(emb_owned.grad, want_g)
        gsum = emb_owned.grad.clone()
        tdist.all_reduce(gsum)
        assert th.allclose(gsum, th.zeros_like(emb).index_add_(0, ids, gsel))
        ret[rank] = "ok"
    except Exception as e:  # surfaced by the parent
        ret[rank] = "FAIL: %r" % (e,)
        raise
    finally:
        tdist.destroy_process_group()


def test_distributed_algorithm_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}


def test_partition_rows_blended_minimises_the_largest_rank_cost():
    """dist.partition_rows_blended: contiguous ranges whose LARGEST cost is minimal (bisection + greedy packing) -
    valid boundaries, never worse than equal-row or equal-edge cuts, exact on a case with a known optimum."""
    from dgl_kgat_b200.dist import partition_rows_blended
    rng = np.random.default_rng(0)
    cost = np.concatenate([np.full(700, 4.0), rng.integers(20, 60, 300).astype(np.float64)])   # cheap user rows, heavy entity rows
    for parts in (2, 4, 8):
        b = partition_rows_blended(cost, parts)
        assert b[0] == 0 and b[-1] == len(cost) and np.all(np.diff(b) >= 0) and len(b) == parts + 1
        worst = max(cost[b[k]: b[k + 1]].sum() for k in range(parts))
        equal_rows = np.linspace(0, len(cost), parts + 1).astype(int)
        assert worst <= max(cost[equal_rows[k]: equal_rows[k + 1]].sum() for k in range(parts)) + 1e-9
        assert worst >= cost.sum() / parts - 1e-9                      # cannot beat the average
        assert worst <= cost.sum() / parts + cost.max() + 1e-9         # greedy packing overshoots by at most one row
    assert partition_rows_blended(np.array([5.0, 1.0, 1.0, 1.0, 1.0, 1.0]), 2).tolist() == [0, 1, 6]
    assert partition_rows_blended(np.ones(5), 1).tolist() == [0, 5]
