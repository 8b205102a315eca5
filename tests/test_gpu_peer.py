"""Peer-memory exchange of the destination-partitioned step (dist.PeerArena, include/kgat_b200.h
section 8) on ONE GPU, so it is covered wherever the GPU tests run:

* world size 1 - the arena, the tensor aliases, the ``*_push`` kernels with no peer, the device-side
  barrier and the CUDA-graph capture of the partitioned step, against the single-GPU step;
* world size 2 with BOTH ranks on GPU 0 - two processes map each other's arena with cudaIpc and run
  the real push / barrier protocol (the GPU time-slices the two contexts, so a spinning barrier kernel
  is preempted for its peer); attention weights bit-identical, embedding / loss / gradients within fp32
  re-association noise of the single-GPU step, and bit-identical ACROSS the two ranks.

The same protocol across two physical GPUs is tests/test_gpu_dist.py (needs >= 2 GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch as th
import torch.distributed as tdist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _single_gpu_reference(m, ck, ids, dev):
    from dgl_kgat_b200.graph import CKGraph
    N, R, trip = ck.cfg.n_nodes, ck.cfg.n_relations, ck.triplets
    g = CKGraph(N, trip, n_rel=R)
    w = m.compute_attention(g)
    g.edata["w"] = w
    emb = m.gnn(g, None)
    loss = m.get_loss(emb, *ids)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    return w, emb.detach(), loss.detach(), grads


def _worker(rank, world, port, ret, skew=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["KGAT_PEER_TIMEOUT_MS"] = "15000"
    # default arithmetic on both sides: the partitioned (PUSH) kernels run the same tensor-core tile GEMMs as the
    # single-GPU ones
    th.cuda.set_device(0)                       # every rank on the same GPU
    dev = th.device("cuda", 0)
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from dgl_kgat_b200 import models
        from dgl_kgat_b200.dist import DistKGAT
        from dgl_kgat_b200.graphed import GraphedDistCFStep
        from dgl_kgat_b200.synth import synth_ckg
        ck = synth_ckg("tiny", skew=skew)       # skew: rows above KGAT_SPLIT_ROW edges inside the partitioned kernels
        N, R, trip = ck.cfg.n_nodes, ck.cfg.n_relations, ck.triplets
        th.manual_seed(1234)
        m = models.Model(True, 64, "kgat", 3, 64, 0.0, n_entities=N, n_relations=R, relation_dim=64,
                         reg_lambda_kg=1e-4, reg_lambda_gnn=1e-4).to(dev)
        rng = np.random.default_rng(0)
        U, I = ck.cfg.n_users, ck.cfg.n_items
        B = 2048
        ids = [th.as_tensor(rng.integers(lo, hi, B), device=dev) for lo, hi in ((0, U), (U, U + I), (U, U + I))]
        w1, emb1, loss1, ref = _single_gpu_reference(m, ck, ids, dev)

        dk = DistKGAT(m, N, trip, R, dev, exchange="peer", max_batch=B)
        assert dk.peer is not None and dk.peer.world == world
        for it in range(3):                     # repeated steps reuse the arena buffers
            m.zero_grad(set_to_none=True)
            dk.compute_attention()
            assert th.equal(w1, dk.attention_weights()), "attention weights differ"
            emb2 = dk.gnn()
            loss2 = dk.get_loss(emb2, *ids)
            loss2.backward()
            r0, r1 = dk.part.r0, dk.part.r1
            scale = emb1.abs().max().item()
            assert (emb1[r0:r1] - emb2[r0:r1]).abs().max().item() <= 1e-6 * scale
            assert abs(loss1.item() - loss2.item()) <= 1e-6 * abs(loss1.item())
            for k, p in m.named_parameters():
                if k in ref:
                    err = (p.grad - ref[k]).abs().max().item() / ref[k].abs().max().item()
                    assert err < 1e-5, (it, k, err)
        eager = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        full = dk.full_embedding(emb2)
        assert (full - emb1).abs().max().item() <= 1e-6 * scale

        # replicated results must be bit-identical on every rank (fixed rank-order sums)
        mine = th.cat([eager[k].reshape(-1) for k in sorted(eager)] + [loss2.detach().reshape(1)]).cpu()
        gathered = [None] * world
        tdist.all_gather_object(gathered, mine)
        for other in gathered:
            assert th.equal(other, gathered[0]), "replicated gradients differ across ranks"

        # the partitioned step as one CUDA graph (no NCCL inside): same numbers.  The eager autograd
        # graph must be gone first: its AccumulateGrad nodes are bound to the default stream.
        loss2_val = loss2.item()
        del emb2, loss2
        m.zero_grad(set_to_none=True)
        step = GraphedDistCFStep(dk, B)
        for _ in range(2):
            loss3 = step(*ids)
        th.cuda.synchronize()
        assert abs(loss3.item() - loss2_val) <= 1e-6 * abs(loss2_val)
        for k, p in m.named_parameters():
            if k in eager:
                assert th.equal(p.grad, eager[k]), "graphed step differs from the eager step: " + k
        assert dk.peer.error() == 0, "a peer barrier timed out"
        step = None

        # the training-loop variant (kgat.py:139-165): attention refreshed per EPOCH outside the graph, the graph
        # holds propagation + loss + backward and must follow every refresh
        m.zero_grad(set_to_none=True)
        dk.compute_attention()
        step2 = GraphedDistCFStep(dk, B, refresh_attention=False)
        step2(*ids)
        th.cuda.synchronize()
        for k, p in m.named_parameters():
            if k in eager:
                assert th.equal(p.grad, eager[k]), "graphed step without refresh differs: " + k
        with th.no_grad():
            m.W_R.mul_(1.05)                    # the next epoch's attention is different
        dk.compute_attention()
        loss5 = step2(*ids)
        th.cuda.synchronize()
        graphed = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
        loss5_val = loss5.item()
        m.zero_grad(set_to_none=True)
        emb6 = dk.gnn()
        loss6 = dk.get_loss(emb6, *ids)
        loss6.backward()
        assert loss6.item() == loss5_val and loss5_val != loss2_val
        for k, p in m.named_parameters():
            if k in graphed:
                assert th.equal(p.grad, graphed[k]), "graphed step did not follow the attention refresh: " + k
        assert dk.peer.error() == 0, "a peer barrier timed out"
        del emb6, loss6
        step2 = None
        dk.close()
        ret[rank] = "ok"
    except Exception as e:
        ret[rank] = "FAIL: %r" % (e,)
        raise
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("world,skew", [(1, False), (2, False), (2, True)])
def test_peer_exchange_on_one_gpu(world, skew):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret, skew), nprocs=world, join=True)
    assert dict(ret) == {r: "ok" for r in range(world)}


def test_push_entry_points_reject_buffers_outside_the_arena():
    """The *_push entry points translate the local output pointer into the peers' arenas; a buffer that
    is not inside the own arena must be refused, not written through a wild pointer."""
    import ctypes
    from dgl_kgat_b200 import _lib
    lib = _lib.load()
    dev = th.device("cuda", 0)
    own, handle = ctypes.c_void_p(), ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
    nbytes = 1 << 20
    with th.cuda.device(dev):
        _lib.check(lib.kgat_peer_arena_create(nbytes, ctypes.byref(own), handle), "create")
        try:
            pg = _lib.KgatPeerGroup()
            pg.world, pg.rank, pg.arena_bytes = 1, 0, nbytes
            pg.arena[0] = own.value
            outside = th.zeros(64, device=dev)
            ids = th.zeros(4, dtype=th.long, device=dev)
            emb = th.zeros(8, 16, device=dev)
            st = _lib.stream_ptr()
            assert lib.kgat_peer_gather_rows_push(emb.data_ptr(), 16, 16, ids.data_ptr(), 4, 0, 8, outside.data_ptr(),
                                                  ctypes.byref(pg), st) == -1
            # inside the header is refused as well; inside the payload is accepted
            assert lib.kgat_peer_gather_rows_push(emb.data_ptr(), 16, 16, ids.data_ptr(), 4, 0, 8, own.value + 512,
                                                  ctypes.byref(pg), st) == -1
            assert lib.kgat_peer_gather_rows_push(emb.data_ptr(), 16, 16, ids.data_ptr(), 4, 0, 8,
                                                  own.value + _lib.PEER_HEADER_BYTES, ctypes.byref(pg), st) == 0
            assert lib.kgat_peer_barrier(ctypes.byref(pg), 1000, st) == 0
            assert lib.kgat_peer_error(ctypes.byref(pg)) == 0
            pg.world = 9
            assert lib.kgat_peer_barrier(ctypes.byref(pg), 1000, st) == -1
        finally:
            th.cuda.synchronize()
            lib.kgat_peer_arena_destroy(own.value)
