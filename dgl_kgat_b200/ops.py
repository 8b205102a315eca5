"""Operator layer: torch tensors in, C-ABI kernel launches out (include/kgat_b200.h).

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); every
number on the hot path is produced by the sm_100a kernels in ``csrc/``.
"""
from __future__ import annotations

import os

import numpy as np
import torch as th

from . import _lib
from .graph import DeviceGraph

LEAKY_SLOPE = 0.01      # F.leaky_relu default used at reference models.py:66
# KGAT_BWD_SPLIT=0 keeps the transposed aggregation inside layer_bwd (A/B measurements)
SPLIT_BWD_GATHER = os.environ.get("KGAT_BWD_SPLIT", "1") != "0"
REDUCE_OVERLAP = os.environ.get("KGAT_BWD_REDUCE_OVERLAP", "1") != "0"      # dW2 reductions on a second stream
_SIDE_STREAMS = {}


def _side_stream(dev):
    s = _SIDE_STREAMS.get(dev)
    if s is None:
        s = _SIDE_STREAMS[dev] = th.cuda.Stream(device=dev)
    return s


def _f32c(t: th.Tensor, device) -> th.Tensor:
    if t.device != device:
        raise _lib.KgatError("tensor on %s, graph structure on %s" % (t.device, device))
    if t.dtype != th.float32:
        raise _lib.KgatError("the KGAT kernels compute in fp32; got %s" % t.dtype)
    return t.contiguous()


def attention_impl_for(n_runs: int, n_edges: int) -> str:
    """The logits kernel the library picks for a graph with ``n_runs`` (head, relation) runs: the run shortcut
    when runs are at most a third of the edges, the per-edge kernel otherwise (KGAT_ATT_IMPL overrides)."""
    impl = os.environ.get("KGAT_ATT_IMPL", "auto")
    if impl != "auto":
        return impl
    return "runs" if 3 * n_runs <= n_edges else "f16"


def attention_plan_for(rel_edges, rel_runs):
    """(impl, rel_mode) from the per-relation edge and (head, relation)-run counts of the WHOLE graph.
    impl: "runs" (the run shortcut) when runs are at most a third of the edges, the per-edge kernel "f16" otherwise.
    rel_mode[r] = 1 where relation r alone passes that test (kgat_runs_build applies the same rule on the device):
    the marks of the per-RELATION choice, KGAT_ATT_IMPL=hybrid.  That one is opt-in: on the synthetic Amazon-book
    graph (random KG heads next to the two interaction relations, whose rows are one run each) it measured
    0.377 ms against 0.344 ms for the per-edge kernel alone - a second tensor-core kernel with its own set-up, W_r
    staging and tail costs more than the 30 % of the projections it saves.
    KGAT_ATT_IMPL = runs / runs_fused / hybrid / f16 / tf32 / ffma forces a kernel (A/B runs)."""
    rel_edges, rel_runs = np.asarray(rel_edges, dtype=np.int64), np.asarray(rel_runs, dtype=np.int64)
    mode = ((rel_edges > 0) & (3 * rel_runs <= rel_edges)).astype(np.int32)
    impl = os.environ.get("KGAT_ATT_IMPL", "auto")
    if impl == "auto":
        impl = "runs" if 3 * int(rel_runs.sum()) <= int(rel_edges.sum()) else "f16"
    return impl, mode


def attention(dg: DeviceGraph, entity: th.Tensor, rel_emb: th.Tensor, W_R: th.Tensor, rows=None, want_eid=True,
              w_csr_out=None, peer=None, impl=None):
    """Logits + per-destination softmax (reference models.py:146-154).

    Returns ``(w_eid, w_csr)``: ``w_eid`` is the (E, 1) tensor ``compute_attention``
    returns (original edge order); ``w_csr`` the same weights in CSR order.
    Multi-GPU (``peer`` = dist.PeerArena): ``w_csr_out`` is this rank's slice of the all-edges weight
    vector inside the arena and the softmax kernel stores every weight to the peers' copies as well.
    """
    lib = _lib.load()
    dev = dg.device
    entity, rel_emb, W_R = (_f32c(t.detach(), dev) for t in (entity, rel_emb, W_R))
    d, rd = entity.shape[1], rel_emb.shape[1]
    if tuple(W_R.shape) != (dg.n_rel, d, rd) or entity.shape[0] != dg.n_nodes or rel_emb.shape[0] != dg.n_rel:
        raise _lib.KgatError("attention: parameter shapes do not match the graph (N=%d, R=%d)" % (dg.n_nodes, dg.n_rel))
    E = max(dg.n_edges, 1)
    logits = th.empty(E, dtype=th.float32, device=dev)
    if w_csr_out is not None:
        if w_csr_out.dtype != th.float32 or w_csr_out.numel() != dg.n_edges or not w_csr_out.is_contiguous():
            raise _lib.KgatError("attention: w_csr_out must be a contiguous fp32 vector of %d weights" % dg.n_edges)
        w_csr = w_csr_out
    else:
        w_csr = th.empty(E, dtype=th.float32, device=dev)
    w_eid = None
    if want_eid:
        w_eid = th.zeros(dg.n_edges, 1, dtype=th.float32, device=dev) if rows is not None else \
            th.empty(dg.n_edges, 1, dtype=th.float32, device=dev)
    r0, r1 = (0, dg.n_nodes) if rows is None else rows
    with th.cuda.device(dev):
        st = _lib.stream_ptr()
        # 64-wide embeddings: pi = e_t . (W_r tanh(e_h W_r + e_r)) needs the two projections once per (head,
        # relation) RUN, not per edge (kgat_attention_logits_runs).  It pays when runs are much fewer than edges -
        # always on a real CKG (every user row is one run), 12.7 x on the Last-FM shape - and not on a KG with
        # uniformly random heads (Amazon-book synthetic: 0.52 runs per edge), where the per-edge kernel stays the
        # only kernel for the KG relations, while its two interaction relations take the shortcut ("hybrid":
        # attention_plan_for).  KGAT_ATT_IMPL = runs / runs_fused / hybrid / f16 / tf32 / ffma forces a kernel.
        # ``impl``: the caller's choice (dist.DistKGAT decides from the GLOBAL run count so that every rank - and the
        # single-GPU step - runs the same arithmetic); None: decide from this graph
        can_runs = d == 64 and rd == 64 and dg.n_rel <= 1024 and dg.n_edges > 0
        if impl is None:
            impl = os.environ.get("KGAT_ATT_IMPL", "auto")
            if impl in ("auto", "hybrid") and can_runs:
                dg.runs()
                impl, _ = attention_plan_for(dg.rel_edges, dg.rel_runs)
        spans = dg.hybrid_spans() if (can_runs and impl == "hybrid") else None     # None: no marked edge on this rank
        use_runs = can_runs and impl in ("runs", "runs_fused")
        runs = dg.runs() if (use_runs or spans is not None) else None     # built once per graph, outside any capture
        with _lib.record("attention_logits", 5 if spans is not None else (4 if use_runs else 3)):
            if spans is not None:
                v_rows = th.empty(max(spans["v_count"], 1), d, dtype=th.float32, device=dev)
                _lib.check(lib.kgat_attention_logits_hybrid(dg.ref, runs, entity.data_ptr(), rel_emb.data_ptr(),
                                                            W_R.data_ptr(), d, rd, logits.data_ptr(), v_rows.data_ptr(),
                                                            spans["v_base"], spans["v_count"], spans["p_begin"],
                                                            spans["p_end"], st), "kgat_attention_logits_hybrid")
            elif use_runs:
                v_rows = None if impl == "runs_fused" else th.empty(max(dg.n_runs, 1), d, dtype=th.float32, device=dev)
                _lib.check(lib.kgat_attention_logits_runs(dg.ref, runs, entity.data_ptr(), rel_emb.data_ptr(),
                                                          W_R.data_ptr(), d, rd, logits.data_ptr(), _lib.ptr(v_rows), st),
                           "kgat_attention_logits_runs")
            else:
                _lib.check(lib.kgat_attention_logits(dg.ref, entity.data_ptr(), rel_emb.data_ptr(), W_R.data_ptr(),
                                                     d, rd, logits.data_ptr(), st), "kgat_attention_logits")
        with _lib.record("segment_softmax"):
            we = w_eid.data_ptr() if (want_eid and dg.n_edges) else None
            if peer is not None:
                _lib.check(lib.kgat_segment_softmax_push(dg.ref, r0, r1, logits.data_ptr(), w_csr.data_ptr(), we,
                                                         peer.ref, st), "kgat_segment_softmax_push")
            else:
                _lib.check(lib.kgat_segment_softmax(dg.ref, r0, r1, logits.data_ptr(), w_csr.data_ptr(), we, st),
                           "kgat_segment_softmax")
    return w_eid, w_csr


def weights_to_csc(dg: DeviceGraph, w_eid: th.Tensor) -> th.Tensor:
    lib = _lib.load()
    w_csc = th.empty(max(dg.n_edges, 1), dtype=th.float32, device=dg.device)
    with th.cuda.device(dg.device):
        with _lib.record("edge_weights_to_csc"):
            _lib.check(lib.kgat_edge_weights_from_eid(dg.ref, w_eid.data_ptr(), None, w_csc.data_ptr(),
                                                      _lib.stream_ptr()), "kgat_edge_weights_from_eid")
    return w_csc


class PropagatePlan:
    """Everything ``Propagate`` needs besides the differentiable inputs.

    Single GPU: one DeviceGraph, all rows.  Multi-GPU (dist.py): ``dg_fwd`` holds the CSR of the
    in-edges of the owned destination rows ``[r0, r1)``, ``dg_bwd`` the CSC of the out-edges of the
    same rows, and ``exchange(t)`` all-gathers the owned rows of a full-size (N, d) buffer in place;
    ``reduce(t)`` sums a small replicated tensor over ranks.  With ``peer`` (a dist.PeerArena) neither
    is called: the exchanged buffers live in the arena, the ``*_push`` kernels store every owned row to
    all peers while they compute, and only a device-side barrier follows."""

    def __init__(self, dg_fwd, w_csr, w_csc, dg_bwd=None, rows=None, exchange=None, reduce=None, peer=None,
                 finish_weights=None):
        self.dg_fwd, self.dg_bwd = dg_fwd, dg_bwd if dg_bwd is not None else dg_fwd
        self.w_csr, self.w_csc = w_csr, w_csc
        # peer exchange: w_csc (weights of the out-edges, heads owned by OTHER ranks) is built lazily, after a
        # barrier that follows the softmax pushes: finish_weights(barrier_done) -> w_csc
        self.finish_weights = finish_weights
        self.r0, self.r1 = (0, dg_fwd.n_nodes) if rows is None else rows
        self.exchange, self.reduce, self.peer = exchange, reduce, peer

    @property
    def distributed(self):
        return self.exchange is not None or self.peer is not None


class GradChannel:
    """Side channel between the fused BPR head and ``Propagate``: the loss gradient touches at most
    3B of the N embedding rows, so ``BprLoss.backward`` parks it here in compact form (rows +
    row_slot map) and hands autograd a zero-stride zero tensor instead of a dense (N, dim) one."""

    def __init__(self):
        self.compact = None         # (rows (3B, dim), row_slot (N,), rows_node (3B,))


def propagate(plan, masks, single_layer, entity, *weights):
    ch = GradChannel()
    out = Propagate.apply(plan, masks, single_layer, ch, entity, *weights)
    if not single_layer:
        out._kgat_channel = ch
    return out


def _is_zero_sentinel(t):
    return t.dim() == 2 and t.stride() == (0, 0)


class Propagate(th.autograd.Function):
    """All KGATConv layers + normalise + concat of ``Model.gnn`` (reference models.py:156-168)
    as one autograd node; ``single_layer=True`` gives the bare ``KGATConv.forward`` output."""

    @staticmethod
    def forward(ctx, plan: PropagatePlan, masks, single_layer, channel, entity, *weights):
        lib = _lib.load()
        dg = plan.dg_fwd
        dev = dg.device
        entity = _f32c(entity, dev)
        weights = [_f32c(w, dev) for w in weights]
        N, L = dg.n_nodes, len(weights)
        if entity.shape[0] != N:
            raise _lib.KgatError("node feature matrix has %d rows, graph has %d nodes" % (entity.shape[0], N))
        dims = [entity.shape[1]] + [w.shape[0] for w in weights]
        for l, w in enumerate(weights):
            if w.shape[1] != dims[l]:
                raise _lib.KgatError("layer %d weight is %s, expected (*, %d)" % (l, tuple(w.shape), dims[l]))
        masks = [None] * L if masks is None else [None if m is None else _f32c(m, dev) for m in masks]
        peer = plan.peer
        # NCCL exchange: rows of other ranks are zero until exchanged; peer push: every row is written by its owner
        alloc = th.zeros if (plan.distributed and peer is None) else th.empty
        if peer is not None:
            peer.generation += 1
        concat = None if single_layer else alloc(N, sum(dims), dtype=th.float32, device=dev)
        cstride = 0 if concat is None else concat.shape[1]
        saved = []
        h, off = entity, dims[0]
        with th.cuda.device(dev):
            st = _lib.stream_ptr()
            for l in range(L):
                d_in, d_out = dims[l], dims[l + 1]
                h_nb = th.empty(N, d_in, dtype=th.float32, device=dev)
                push = peer is not None and l + 1 < L       # the next layer gathers rows owned by other ranks
                h_out = peer.buf(("h", l), (N, d_out)) if push else alloc(N, d_out, dtype=th.float32, device=dev)
                act = th.empty(N, d_out, dtype=th.float32, device=dev) if masks[l] is not None else None
                with _lib.record("layer_fwd_%dx%d" % (d_in, d_out)):
                    args = (dg.ref, plan.r0, plan.r1, h.data_ptr(), plan.w_csr.data_ptr(), weights[l].data_ptr(),
                            _lib.ptr(masks[l]), d_in, d_out, LEAKY_SLOPE, h_nb.data_ptr(), _lib.ptr(act),
                            h_out.data_ptr(), _lib.ptr(concat), cstride, off,
                            0 if (l == 0 and concat is not None) else -1)
                    if push:
                        _lib.check(lib.kgat_layer_fwd_push(*args, peer.ref, st), "kgat_layer_fwd_push")
                    else:
                        _lib.check(lib.kgat_layer_fwd(*args, st), "kgat_layer_fwd")
                if push:
                    peer.barrier()
                    if plan.w_csc is None:          # this barrier also covered the peers' attention weights
                        plan.w_csc = plan.finish_weights(True)
                elif plan.distributed and l + 1 < L:
                    plan.exchange(h_out)
                saved.append((h, h_nb, act, h_out))
                h, off = h_out, off + d_out
        if plan.w_csc is None and plan.finish_weights is not None:      # single-layer model: no row exchange above
            plan.w_csc = plan.finish_weights(False)
        ctx.plan, ctx.masks, ctx.dims, ctx.single = plan, masks, dims, single_layer
        ctx.saved, ctx.weights, ctx.channel = saved, weights, channel
        ctx.generation = peer.generation if peer is not None else 0
        return h if single_layer else concat

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        plan, dims, L = ctx.plan, ctx.dims, len(ctx.weights)
        dg = plan.dg_bwd
        dev, N = dg.device, dg.n_nodes
        n_part = lib.kgat_layer_bwd_partials()
        dc_slot = None
        compact, ctx.channel.compact = ctx.channel.compact, None
        if compact is not None and not ctx.single:
            rows, row_slot, rows_node = compact
            if _is_zero_sentinel(grad_out):
                d_concat, dc_slot, cstride = rows, row_slot, rows.shape[1]
            else:                       # another consumer contributed a dense gradient as well: densify
                valid = rows_node >= 0
                grad_out = _f32c(grad_out, dev).clone()
                grad_out.index_add_(0, rows_node[valid].long(), rows[valid])
                d_concat, cstride = grad_out, grad_out.shape[1]
        else:
            grad_out = _f32c(grad_out, dev)
            d_concat = None if ctx.single else grad_out
            cstride = 0 if ctx.single else grad_out.shape[1]
        offs = [0]
        for d in dims:
            offs.append(offs[-1] + d)
        d_weights = [None] * L
        nxt_dx, nxt_dh = None, None
        peer = plan.peer
        if peer is not None and ctx.generation != peer.generation:
            raise _lib.KgatError("the exchanged activations of this forward pass were overwritten by a later "
                                 "gnn() call on the same DistKGAT; run backward before the next forward")
        alloc = th.zeros if (plan.distributed and peer is None) else th.empty
        w_off = [0]
        for l in range(L):
            w_off.append(w_off[-1] + dims[l] * dims[l + 1])
        with th.cuda.device(dev):
            st = _lib.stream_ptr()
            # peer exchange: every layer's dW2 is a view of one flat buffer that travels as one slot
            dw_flat = th.empty(w_off[-1], dtype=th.float32, device=dev) if peer is not None else None
            # (single GPU and the peer exchange; the NCCL exchange keeps everything on one stream)
            side = _side_stream(dev) if (REDUCE_OVERLAP and (peer is not None or not plan.distributed)) else None
            keep = []            # partial buffers stay alive until the side stream has been joined
            for l in reversed(range(L)):
                d_in, d_out = dims[l], dims[l + 1]
                h_in, h_nb, act, h_out = ctx.saved[l]
                dx = th.empty(N, d_in, dtype=th.float32, device=dev)
                dh = peer.buf(("dh", l), (N, d_in)) if peer is not None else alloc(N, d_in, dtype=th.float32, device=dev)
                dW = dw_flat[w_off[l]: w_off[l + 1]].view(d_out, d_in) if peer is not None else \
                    th.empty(d_out, d_in, dtype=th.float32, device=dev)
                part = th.empty(n_part * d_out * d_in, dtype=th.float32, device=dev)
                if ctx.single:
                    # bare KGATConv: the incoming gradient IS the gradient of the layer output
                    up_dx, up_dh = grad_out, None
                elif nxt_dh is not None and SPLIT_BWD_GATHER:
                    # the next layer's contribution to this layer's output gradient - its direct term plus the
                    # transposed aggregation of its dh_nb - as a gather kernel of its own; layer_bwd then runs
                    # its dense part only (nxt_dh_nb = NULL)
                    gsum = th.empty(N, d_out, dtype=th.float32, device=dev)
                    with _lib.record("csc_gather_%d" % d_out):
                        _lib.check(lib.kgat_csc_gather(dg.ref, plan.r0, plan.r1, nxt_dx.data_ptr(), nxt_dh.data_ptr(),
                                                       plan.w_csc.data_ptr(), d_out, gsum.data_ptr(), st),
                                   "kgat_csc_gather")
                    up_dx, up_dh = gsum, None
                else:
                    up_dx, up_dh = nxt_dx, nxt_dh
                with _lib.record("layer_bwd_%dx%d" % (d_in, d_out), 1 if side is not None else 2):
                    args = (dg.ref, plan.r0, plan.r1, _lib.ptr(d_concat), _lib.ptr(dc_slot), cstride, offs[l + 1],
                            _lib.ptr(up_dx), _lib.ptr(up_dh), plan.w_csc.data_ptr(),
                            h_in.data_ptr(), h_nb.data_ptr(), _lib.ptr(act), h_out.data_ptr(), _lib.ptr(ctx.masks[l]),
                            ctx.weights[l].data_ptr(), d_in, d_out, LEAKY_SLOPE,
                            dx.data_ptr(), dh.data_ptr(), None if side is not None else dW.data_ptr(), part.data_ptr())
                    if peer is not None:
                        _lib.check(lib.kgat_layer_bwd_push(*args, peer.ref, st), "kgat_layer_bwd_push")
                    else:
                        _lib.check(lib.kgat_layer_bwd(*args, st), "kgat_layer_bwd")
                if side is not None:
                    # nothing in the rest of the backward pass reads dW2: its fixed-order reduction over the per-CTA
                    # partials (a 16..128-CTA kernel, ~10 us of latency) runs on a second stream next to the
                    # following gather kernels and is joined before the gradients are handed back
                    side.wait_stream(th.cuda.current_stream())
                    with th.cuda.stream(side):
                        with _lib.record("dW2_reduce", 1):
                            _lib.check(lib.kgat_layer_bwd_reduce(plan.r1 - plan.r0, d_in, d_out, part.data_ptr(),
                                                                 dW.data_ptr(), _lib.stream_ptr()),
                                       "kgat_layer_bwd_reduce")
                    keep.append(part)
                # dh_nb rows of every rank feed the transposed gather of the next kernel
                if peer is not None:
                    peer.barrier()
                elif plan.distributed:
                    plan.exchange(dh)
                d_weights[l] = dW
                nxt_dx, nxt_dh = dx, dh
            d_entity = None
            if peer is not None:
                # last phase: the dW2 of all layers (one slot per rank) and the entity gradient share a barrier
                slots = peer.buf(("dW2",), (peer.world, w_off[-1]))
                if side is not None:
                    th.cuda.current_stream().wait_stream(side)      # the dW2 of every layer is complete
                with _lib.record("peer_push_slot"):
                    _lib.check(lib.kgat_peer_push_slot(dw_flat.data_ptr(), w_off[-1], slots.data_ptr(), peer.ref, st),
                               "kgat_peer_push_slot")
            if ctx.needs_input_grad[4]:
                d_entity = peer.buf(("d_entity",), (N, dims[0])) if peer is not None else \
                    alloc(N, dims[0], dtype=th.float32, device=dev)
                with _lib.record("input_bwd_%d" % dims[0]):
                    args = (dg.ref, plan.r0, plan.r1, _lib.ptr(d_concat), _lib.ptr(dc_slot), cstride, 0,
                            nxt_dx.data_ptr(), nxt_dh.data_ptr(), plan.w_csc.data_ptr(), dims[0], d_entity.data_ptr())
                    if peer is not None:
                        _lib.check(lib.kgat_input_bwd_push(*args, peer.ref, st), "kgat_input_bwd_push")
                    else:
                        _lib.check(lib.kgat_input_bwd(*args, st), "kgat_input_bwd")
                if peer is None and plan.distributed:
                    plan.exchange(d_entity)             # replicated parameters: every rank needs the full gradient
            if peer is not None:
                peer.barrier()
                dw_sum = th.empty(w_off[-1], dtype=th.float32, device=dev)
                with _lib.record("peer_sum_slots"):
                    _lib.check(lib.kgat_peer_sum_slots(slots.data_ptr(), w_off[-1], peer.world, dw_sum.data_ptr(), st),
                               "kgat_peer_sum_slots")
                d_weights = [dw_sum[w_off[l]: w_off[l + 1]].view(dims[l + 1], dims[l]) for l in range(L)]
                if d_entity is not None:
                    d_entity = d_entity.clone()         # the arena buffer is overwritten by the next backward
            elif plan.distributed:
                plan.reduce(d_weights)                  # one all-reduce for every layer's dW2
            if side is not None:
                th.cuda.current_stream().wait_stream(side)
        return (None, None, None, None, d_entity, *d_weights)


def _bpr_forward(table, ids, reg_lambda):
    """Launch the fused forward on ``table`` rows addressed by the three int64 id vectors."""
    lib = _lib.load()
    dev = table.device
    B, dim = ids[0].numel(), table.shape[1]
    ws_bytes = lib.kgat_bpr_workspace_bytes(B)
    ws = th.empty(ws_bytes, dtype=th.uint8, device=dev)
    loss = th.empty((), dtype=th.float32, device=dev)
    sig = th.empty(B, dtype=th.float32, device=dev)
    with th.cuda.device(dev):
        with _lib.record("bpr_loss_fwd", 2):
            _lib.check(lib.kgat_bpr_loss_fwd(table.data_ptr(), table.stride(0), dim, ids[0].data_ptr(), ids[1].data_ptr(),
                                             ids[2].data_ptr(), B, float(reg_lambda), loss.data_ptr(), sig.data_ptr(),
                                             ws.data_ptr(), ws_bytes, _lib.stream_ptr()), "kgat_bpr_loss_fwd")
    return loss, sig, ws


def _bpr_backward(table, ids, node_ids, own, n_nodes, reg_lambda, sig, grad_loss, ws):
    """Compact gradient (rows, rows_node, row_slot) w.r.t. the (n_nodes, dim) embedding."""
    lib = _lib.load()
    dev = table.device
    B, dim = ids[0].numel(), table.shape[1]
    g = grad_loss.to(device=dev, dtype=th.float32).contiguous()
    rows = th.empty(3 * B, dim, dtype=th.float32, device=dev)
    rows_node = th.empty(3 * B, dtype=th.int32, device=dev)
    row_slot = th.empty(n_nodes, dtype=th.int32, device=dev)
    nptr = [None] * 3 if node_ids is None else [t.data_ptr() for t in node_ids]
    with th.cuda.device(dev):
        with _lib.record("bpr_loss_bwd", 3):
            _lib.check(lib.kgat_bpr_loss_bwd(table.data_ptr(), table.stride(0), dim, ids[0].data_ptr(), ids[1].data_ptr(),
                                             ids[2].data_ptr(), B, float(reg_lambda), sig.data_ptr(), g.data_ptr(),
                                             nptr[0], nptr[1], nptr[2], own[0], own[1], n_nodes,
                                             rows.data_ptr(), rows_node.data_ptr(), row_slot.data_ptr(),
                                             ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "kgat_bpr_loss_bwd")
    return rows, rows_node, row_slot


def _deliver_compact(channel, compact, shape, dev):
    """Park the compact gradient on the channel (zero-stride sentinel for autograd) or densify it."""
    rows, rows_node, row_slot = compact
    if channel is not None and channel.compact is None:
        channel.compact = (rows, row_slot, rows_node)
        return th.zeros(1, 1, dtype=th.float32, device=dev).expand(*shape)
    valid = rows_node >= 0
    grad = th.zeros(*shape, dtype=th.float32, device=dev)
    grad.index_add_(0, rows_node[valid].long(), rows[valid])
    return grad


class BprLoss(th.autograd.Function):
    """Fused BPR loss head (reference models.py:170-178): forward = gather + dots + logsigmoid +
    L2 terms in one kernel; backward = compact, deterministic row gradient (see GradChannel)."""

    @staticmethod
    def forward(ctx, embedding, src, pos, neg, reg_lambda, channel):
        dev = embedding.device
        emb = _f32c(embedding, dev)
        ids = [i.to(device=dev, dtype=th.int64).contiguous().view(-1) for i in (src, pos, neg)]
        if ids[1].numel() != ids[0].numel() or ids[2].numel() != ids[0].numel():
            raise _lib.KgatError("get_loss: id vectors differ in length")
        loss, sig, ws = _bpr_forward(emb, ids, reg_lambda)
        ctx.emb, ctx.ids, ctx.sig, ctx.lam, ctx.channel, ctx.ws = emb, ids, sig, float(reg_lambda), channel, ws
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        emb = ctx.emb
        N = emb.shape[0]
        compact = _bpr_backward(emb, ctx.ids, None, (0, N), N, ctx.lam, ctx.sig, grad_loss, ctx.ws)
        return _deliver_compact(ctx.channel, compact, emb.shape, emb.device), None, None, None, None, None


class DistBprLoss(th.autograd.Function):
    """The same loss on a destination-partitioned embedding: every batch row is contributed by its
    owner and summed over ranks (one all-reduce of the (3B, dim) table), the loss is replicated, and the
    backward needs no communication - each rank keeps the gradient rows of the nodes it owns."""

    @staticmethod
    def forward(ctx, embedding, src, pos, neg, reg_lambda, channel, own, group, peer=None):
        import torch.distributed as tdist
        dev = embedding.device
        emb = _f32c(embedding, dev)
        node_ids = [i.to(device=dev, dtype=th.int64).contiguous().view(-1) for i in (src, pos, neg)]
        B = node_ids[0].numel()
        cat = th.cat(node_ids)
        if peer is not None:
            # every owner stores its rows into the table of every rank (NVLink stores), then one barrier
            # ONE table sized for the largest batch (DistKGAT(max_batch=...)): the partial last batch of an epoch
            # takes a prefix of it instead of a second arena buffer
            full = peer._bufs.get(("loss_rows", emb.shape[1]))
            cap = full.shape[0] // 3 if full is not None else max(int(getattr(peer, "max_batch", 0)), B)
            if B > cap:
                raise _lib.KgatError("batch of %d rows exceeds the %d the peer arena was sized for (DistKGAT max_batch)"
                                     % (B, cap))
            table = peer.buf(("loss_rows", emb.shape[1]), (3 * cap, emb.shape[1]))[: 3 * B]
            with th.cuda.device(dev):
                with _lib.record("peer_gather_rows_push"):
                    _lib.check(_lib.load().kgat_peer_gather_rows_push(
                        emb.data_ptr(), emb.stride(0), emb.shape[1], cat.data_ptr(), 3 * B, own[0], own[1],
                        table.data_ptr(), peer.ref, _lib.stream_ptr()), "kgat_peer_gather_rows_push")
            peer.barrier()
        else:
            mine = ((cat >= own[0]) & (cat < own[1])).unsqueeze(1)
            table = emb.index_select(0, cat) * mine
            tdist.all_reduce(table, group=group)
        ar = th.arange(B, device=dev)
        ids = [ar, ar + B, ar + 2 * B]
        loss, sig, ws = _bpr_forward(table, ids, reg_lambda)
        ctx.table, ctx.ids, ctx.node_ids, ctx.sig, ctx.ws = table, ids, node_ids, sig, ws
        ctx.lam, ctx.channel, ctx.own, ctx.shape = float(reg_lambda), channel, own, emb.shape
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        compact = _bpr_backward(ctx.table, ctx.ids, ctx.node_ids, ctx.own, ctx.shape[0], ctx.lam, ctx.sig,
                                grad_loss, ctx.ws)
        grad = _deliver_compact(ctx.channel, compact, ctx.shape, ctx.table.device)
        return grad, None, None, None, None, None, None, None, None


def bpr_loss(embedding, src, pos, neg, reg_lambda):
    return BprLoss.apply(embedding, src, pos, neg, reg_lambda, getattr(embedding, "_kgat_channel", None))


def score_topk(emb: th.Tensor, user_ids: th.Tensor, item_ids: th.Tensor, mask_ptr, mask_items, K: int,
               want_scores: bool = False):
    """Fused scoring + per-row top-K (reference metric.py:44-51).  ids are int32 device tensors;
    returns (n_users, K) int32 positions into ``item_ids`` (and the scores when asked)."""
    lib = _lib.load()
    dev = emb.device
    emb = _f32c(emb.detach(), dev)
    n_users, n_items = user_ids.numel(), item_ids.numel()
    if n_items < K:
        raise _lib.KgatError("top-%d of %d items" % (K, n_items))
    topk = th.empty(n_users, K, dtype=th.int32, device=dev)
    scores = th.empty(n_users, K, dtype=th.float32, device=dev) if want_scores else None
    # the pre-split operand tiles and the per-range top-K lists of the fast path live in a workspace
    ws_bytes = lib.kgat_score_topk_workspace_bytes(n_users, n_items, emb.shape[1], K)
    ws = th.empty(max(ws_bytes, 16), dtype=th.uint8, device=dev)
    with th.cuda.device(dev):
        with _lib.record("score_topk", 4):
            _lib.check(lib.kgat_score_topk_ws(emb.data_ptr(), emb.stride(0), emb.shape[1], user_ids.data_ptr(), n_users,
                                              item_ids.data_ptr(), n_items, _lib.ptr(mask_ptr),
                                              _lib.ptr(mask_items) if mask_items is not None and mask_items.numel() else None,
                                              K, topk.data_ptr(), _lib.ptr(scores), ws.data_ptr(), ws_bytes,
                                              _lib.stream_ptr()), "kgat_score_topk_ws")
    return (topk, scores) if want_scores else topk


# --------------------------------------------------------------------------- TransR step
def transr_supported(entity: th.Tensor, rel_emb: th.Tensor, W_R: th.Tensor, batch: int) -> bool:
    """Shapes the TransR kernels cover (Model.transR refuses anything else on the GPU)."""
    return (entity.is_cuda and entity.dtype == th.float32 and entity.shape[1] == 64 and rel_emb.shape[1] == 64
            and W_R.shape[0] <= 1024 and batch > 0)


class TransRLoss(th.autograd.Function):
    """Model.transR (reference models.py:114-133): ``forward`` runs the relation-grouped pass that
    yields the loss and leaves every gradient piece in the workspace; ``backward`` turns them into the
    three dense gradients, scaled by the upstream gradient on the device."""

    @staticmethod
    def forward(ctx, entity, rel_emb, W_R, h, r, pos_t, neg_t, reg_lambda):
        lib = _lib.load()
        dev = entity.device
        ent, rel, W = (_f32c(t.detach(), dev) for t in (entity, rel_emb, W_R))
        ids = [t.to(device=dev, dtype=th.long).contiguous() for t in (h, r, pos_t, neg_t)]
        B, R = ids[0].numel(), rel.shape[0]
        ws_bytes = lib.kgat_transr_workspace_bytes(B, R)
        ws = th.empty(ws_bytes, dtype=th.uint8, device=dev)
        loss = th.empty((), dtype=th.float32, device=dev)
        with th.cuda.device(dev):
            with _lib.record("transr_fwd", 2):
                _lib.check(lib.kgat_transr_fwd(ent.data_ptr(), ent.shape[0], rel.data_ptr(), W.data_ptr(), R,
                                               ent.shape[1], rel.shape[1], ids[0].data_ptr(), ids[1].data_ptr(),
                                               ids[2].data_ptr(), ids[3].data_ptr(), B, float(reg_lambda),
                                               loss.data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()),
                           "kgat_transr_fwd")
        ctx.save_for_backward(ws)
        ctx.shapes = (tuple(ent.shape), tuple(rel.shape), tuple(W.shape), B)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        lib = _lib.load()
        (ws,) = ctx.saved_tensors
        es, rs, Ws_, B = ctx.shapes
        dev = ws.device
        g = grad_loss.to(device=dev, dtype=th.float32).contiguous()
        d_ent = th.empty(es, dtype=th.float32, device=dev)
        d_rel = th.empty(rs, dtype=th.float32, device=dev)
        d_W = th.empty(Ws_, dtype=th.float32, device=dev)
        with th.cuda.device(dev):
            with _lib.record("transr_bwd", 3):
                _lib.check(lib.kgat_transr_bwd(g.data_ptr(), es[0], rs[0], es[1], rs[1], B, d_ent.data_ptr(),
                                               d_rel.data_ptr(), d_W.data_ptr(), ws.data_ptr(), ws.numel(),
                                               _lib.stream_ptr()), "kgat_transr_bwd")
        return d_ent, d_rel, d_W, None, None, None, None, None


TRANSR_MAX_BATCH = 16384 // 3          # (node, row) keys one plan kernel sorts


def transr_loss(entity, rel_emb, W_R, h, r, pos_t, neg_t, reg_lambda):
    """Every term of the loss is a mean over the batch, so a batch above the plan kernel's capacity is the
    size-weighted sum of its chunks (``--batch_size_kg`` beyond 5461)."""
    B = h.numel()
    if B <= TRANSR_MAX_BATCH:
        return TransRLoss.apply(entity, rel_emb, W_R, h, r, pos_t, neg_t, reg_lambda)
    total = None
    for b0 in range(0, B, TRANSR_MAX_BATCH):
        b1 = min(B, b0 + TRANSR_MAX_BATCH)
        part = TransRLoss.apply(entity, rel_emb, W_R, h[b0:b1], r[b0:b1], pos_t[b0:b1], neg_t[b0:b1], reg_lambda)
        part = part * ((b1 - b0) / B)
        total = part if total is None else total + part
    return total
