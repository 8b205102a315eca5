"""Multi-GPU KGAT: 1-D partition by destination entity, one process per GPU.

New capability (the reference is single-GPU, kgat.py:14,69-71).  Every rank holds the replicated
parameters, the CSR of the in-edges of its own contiguous destination range (forward, attention)
and the CSC of the out-edges of the same range (backward).  Ranges are balanced by edge count.

Two exchange implementations (``DistKGAT(..., exchange=...)``, default from ``KGAT_DIST_EXCHANGE``):

``"peer"``  every rank owns an arena that all other ranks map over NVLink / NVSwitch (cudaIpc).  The
            kernels that PRODUCE exchanged rows (softmax, layer forward, layer backward, entity
            gradient, loss-row gather) store each owned row into the own arena and into every peer's
            arena in the same instruction stream, so the all-gather rides on the compute kernel; what
            remains of each collective is one device-side barrier (``kgat_peer_barrier``, ~5 us, no host,
            no NCCL).  8 barriers per step, the whole step is CUDA-graph capturable
            (``graphed.GraphedDistCFStep``).  ``torch.distributed`` only carries the 64-byte IPC handles.
``"nccl"``  the same algorithm with NCCL collectives between unchanged single-GPU kernels:
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

import ctypes
import os
import socket

from . import _lib, metric, ops
from .graph import CKGraph


class _RawCuda:
    """``__cuda_array_interface__`` view of a piece of device memory (fp32) for ``torch.as_tensor``."""

    def __init__(self, ptr, shape, owner):
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f4",
                                         "data": (int(ptr), False), "version": 2, "strides": None}


class PeerArena:
    """This rank's arena plus the mapped arenas of all peers (include/kgat_b200.h section 8).

    ``buf(name, shape)`` carves a persistent fp32 tensor out of the own arena; every rank must request
    the same names and shapes in the same order so a buffer has the SAME offset in every arena - that
    offset is how the ``*_push`` kernels find the peers' copies.  ``barrier()`` enqueues the device-side
    barrier on the current stream; every rank must enqueue the same sequence."""

    def __init__(self, payload_bytes, device, group=None, timeout_ms=0):
        lib = _lib.load()
        self.device, self.group = th.device(device), group
        self.rank, self.world = tdist.get_rank(group), tdist.get_world_size(group)
        if self.world > _lib.MAX_RANKS:
            raise _lib.KgatError("peer exchange supports at most %d ranks (one NVLink box)" % _lib.MAX_RANKS)
        self.nbytes = (_lib.PEER_HEADER_BYTES + int(payload_bytes) + 255) // 256 * 256
        self.timeout_ms = int(timeout_ms) or int(os.environ.get("KGAT_PEER_TIMEOUT_MS", "0"))
        self.generation = 0
        self._bufs, self._top, self._closed = {}, _lib.PEER_HEADER_BYTES, False
        self._pg = _lib.KgatPeerGroup()
        self._pg.world, self._pg.rank, self._pg.arena_bytes = self.world, self.rank, self.nbytes
        self._pg.multicast = None
        self._mapped, self._symm = [], None
        # NVSwitch multicast pays from three ranks on (measured, profiles/README.md: 0.71 vs 0.80 ms per step at 4
        # ranks); at two ranks a unicast store already leaves the GPU once, and the multicast would send the own
        # copy through the switch and back on top (0.98 vs 0.91 ms).  KGAT_PEER_MULTICAST=1 / 0 forces it.
        mc_env = os.environ.get("KGAT_PEER_MULTICAST", "auto")
        want_mc = self.world > 1 and (mc_env == "1" or (mc_env not in ("0", "1") and self.world > 2))
        with th.cuda.device(self.device):
            if want_mc and self._try_symmetric(group):
                pass
            else:
                self._ipc_arena(lib, group)
            self.ref = ctypes.byref(self._pg)
            tdist.barrier(group=group)       # nobody pushes before every rank has mapped every arena

    def _try_symmetric(self, group):
        """Arena from torch's symmetric memory: the rendezvous maps every rank's allocation into this process AND
        binds them to one NVSwitch multicast object, whose address lets a single ``multimem.st`` reach all ranks
        (include/kgat_b200.h, kgat_peer_group_t.multicast).  torch is the plumbing here (VMM allocation, handle
        exchange); returns False when this build / fabric offers no multicast, and the cudaIpc arena is used."""
        t = h = None
        mc, ptrs = 0, []
        try:
            import torch.distributed._symmetric_memory as symm
            g = group if group is not None else tdist.group.WORLD
            t = symm.empty(self.nbytes, dtype=th.uint8, device=self.device)
            t.zero_()
            th.cuda.synchronize()
            h = symm.rendezvous(t, g.group_name)
            mc = int(getattr(h, "multicast_ptr", 0) or 0)
            ptrs = [int(x) for x in h.buffer_ptrs]
        except Exception:            # no symmetric memory in this torch / no fabric support: unicast pushes
            mc = 0
        # every rank reaches this vote, whatever happened above
        ok = th.tensor([1 if (mc != 0 and len(ptrs) == self.world and ptrs[self.rank] == t.data_ptr()) else 0],
                       device=self.device)
        tdist.all_reduce(ok, op=tdist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            return False
        self._symm = (t, h)
        self._own = t.data_ptr()
        for k in range(self.world):
            self._pg.arena[k] = ptrs[k]
        self._pg.multicast = mc
        return True

    def _ipc_arena(self, lib, group):
        own = ctypes.c_void_p()
        handle = ctypes.create_string_buffer(_lib.IPC_HANDLE_BYTES)
        _lib.check(lib.kgat_peer_arena_create(self.nbytes, ctypes.byref(own), handle), "kgat_peer_arena_create")
        self._own = own.value
        infos = [None] * self.world
        tdist.all_gather_object(infos, (bytes(handle.raw), self.nbytes, socket.gethostname()), group=group)
        if len({host for _, _, host in infos}) != 1:
            raise _lib.KgatError("the peer exchange maps the other ranks' memory with cudaIpc and needs all ranks "
                                 "on one NVLink box; ranks span %s - use exchange='nccl'"
                                 % sorted({host for _, _, host in infos}))
        if any(n != self.nbytes for _, n, _ in infos):
            raise _lib.KgatError("peer arenas differ in size across ranks: %s" % [n for _, n, _ in infos])
        for k, (h, _, _) in enumerate(infos):
            if k == self.rank:
                self._pg.arena[k] = self._own
                continue
            p = ctypes.c_void_p()
            _lib.check(lib.kgat_peer_arena_open(h, ctypes.byref(p)), "kgat_peer_arena_open (rank %d)" % k)
            self._pg.arena[k] = p.value
            self._mapped.append(p.value)

    @property
    def multicast(self):
        return bool(self._pg.multicast)

    def buf(self, name, shape):
        t = self._bufs.get(name)
        shape = tuple(int(x) for x in shape)
        if t is not None:
            if tuple(t.shape) != shape:
                raise _lib.KgatError("arena buffer %r has shape %s, requested %s" % (name, tuple(t.shape), shape))
            return t
        n = 4
        for x in shape:
            n *= x
        off = self._top
        if off + n > self.nbytes:
            raise _lib.KgatError("peer arena exhausted (%d bytes): buffer %r needs %d more; raise max_batch / "
                                 "arena size in DistKGAT" % (self.nbytes, name, n))
        self._top = (off + n + 255) // 256 * 256
        t = th.as_tensor(_RawCuda(self._own + off, shape, self), device=self.device)
        if t.data_ptr() != self._own + off or t.dtype != th.float32:
            raise _lib.KgatError("could not alias arena memory as a torch tensor")
        self._bufs[name] = t
        return t

    def barrier(self):
        with th.cuda.device(self.device):
            with _lib.record("peer_barrier"):
                _lib.check(_lib.load().kgat_peer_barrier(self.ref, self.timeout_ms, _lib.stream_ptr()),
                           "kgat_peer_barrier")

    def error(self):
        """0, or 1 + the rank that did not reach a barrier in time (synchronises the device)."""
        with th.cuda.device(self.device):
            return int(_lib.load().kgat_peer_error(self.ref))

    def check(self):
        """Raise if a barrier of this rank ever timed out.  After a timeout the barriers stop waiting (so the
        GPU does not hang), which means every later phase reads half-pushed rows: the training / eval drivers
        call this at their synchronisation points so a lost peer is an error, not silently corrupted numbers."""
        e = self.error()
        if e:
            raise _lib.KgatError("peer barrier timed out on rank %d waiting for rank %d: the exchanged buffers "
                                 "are incomplete, results since then are invalid" % (self.rank, e - 1))

    def close(self):
        """Collective: unmap the peers' arenas, then free the own one."""
        if self._closed:
            return
        self._closed = True
        lib = _lib.load()
        with th.cuda.device(self.device):
            th.cuda.synchronize()
            tdist.barrier(group=self.group)
            self._bufs.clear()
            if self._symm is not None:
                tdist.barrier(group=self.group)
                self._symm = None                   # the symmetric allocation is released with its tensor
                return
            for p in self._mapped:
                lib.kgat_peer_arena_close(p)
            tdist.barrier(group=self.group)
            lib.kgat_peer_arena_destroy(self._own)


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


def partition_rows_minmax(edge_cost: np.ndarray, n_parts: int, row_ns: float, edge_ns: float) -> np.ndarray:
    """Boundaries b[0..n_parts] of contiguous ranges minimising
    ``row_ns * max_k rows(k) + edge_ns * max_k edges(k)``.

    The phases of a partitioned step are separated by barriers, and a phase is bound either by the ROWS a
    rank owns (the kernels that push every owned row to all peers over NVLink) or by its EDGES (the
    gathers), so the step costs the sum of the per-phase maxima, not the maximum of the per-rank sums:
    balancing a blended per-row cost leaves one rank row-heavy and the others edge-heavy and every phase
    waits for somebody.  For a row budget R and an edge budget E the greedy left-to-right packing decides
    feasibility exactly (both constraints are monotone in the range), so the optimum is found by scanning
    R and bisecting E.  Deterministic: every rank computes the same boundaries."""
    n = len(edge_cost)
    if n_parts <= 1 or n == 0:
        return np.array([0] + [n] * max(n_parts, 1), dtype=np.int64)
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(edge_cost, out=ptr[1:])
    total, heaviest = int(ptr[-1]), int(edge_cost.max())

    def pack(r_max, e_max):
        b = [0]
        for _ in range(n_parts):
            a = b[-1]
            by_edges = int(np.searchsorted(ptr, ptr[a] + e_max, side="right")) - 1
            b.append(max(a, min(n, a + r_max, by_edges)))
        return b if b[-1] == n else None

    best = None
    r_lo = -(-n // n_parts)
    for r_max in sorted({int(r) for r in np.linspace(r_lo, n, 48)}):
        lo, hi = max(heaviest, -(-total // n_parts)), total
        if pack(r_max, hi) is None:
            continue
        while lo < hi:
            mid = (lo + hi) // 2
            if pack(r_max, mid) is None:
                lo = mid + 1
            else:
                hi = mid
        cost = row_ns * r_max + edge_ns * lo
        if best is None or cost < best[0]:
            best = (cost, pack(r_max, lo))
    return np.asarray(best[1], dtype=np.int64)


def partition_rows_blended(row_cost: np.ndarray, n_parts: int) -> np.ndarray:
    """Boundaries b[0..n_parts] of contiguous ranges minimising the LARGEST range cost, cost = sum of ``row_cost``
    over the range (bisection on the bound, greedy left-to-right packing decides feasibility exactly).

    Measured on 4 B200s (profiles/README.md): the step time of the partitioned step is the kernel time of its
    slowest rank, and a rank's kernel time is linear in its rows and edges (12 ns per row with unicast pushes to 3
    peers, 0.29 ns per in-edge) - minimising the per-phase maxima separately (partition_rows_minmax) left one
    rank with both the most rows and the most edges."""
    n = len(row_cost)
    if n_parts <= 1 or n == 0:
        return np.array([0] + [n] * max(n_parts, 1), dtype=np.int64)
    ptr = np.zeros(n + 1, dtype=np.float64)
    np.cumsum(row_cost, out=ptr[1:])
    total, heaviest = float(ptr[-1]), float(row_cost.max())

    def pack(bound):
        b = [0]
        for _ in range(n_parts):
            a = b[-1]
            b.append(max(a, int(np.searchsorted(ptr, ptr[a] + bound, side="right")) - 1))
        return b if b[-1] >= n else None

    lo, hi = max(heaviest, total / n_parts), total
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if pack(mid) is None:
            lo = mid
        else:
            hi = mid
    b = pack(hi)
    b[-1] = n
    return np.asarray(b, dtype=np.int64)


class Partition:
    """The slice of the CKG owned by ``rank``: edge subsets and the two local graphs.

    ``row_ns`` = 0 (default): contiguous ranges with equal in-edge counts (``partition_rows``).
    ``row_ns`` > 0: ranges that minimise ``row_ns * max rows + edge_ns * max edges`` over the ranks
    (``partition_rows_minmax``; an edge = half an in-edge + half an out-edge, forward and backward gathers).
    Balancing edges alone gives the rank that holds the users (~8 edges each, 44 % of the rows of an
    Amazon-book CKG) half of all rows at 4 ranks - measured: its push kernels were NVLink-egress bound
    (1.7 ns per 256-byte row and 3 peers) while the other ranks idled at the barriers."""

    def __init__(self, n_nodes, triplets, n_rel, rank, world, row_ns=0.0, edge_ns=0.36, blended=False):
        trip = np.ascontiguousarray(triplets, dtype=np.int32)
        self.n_nodes, self.n_rel, self.n_edges = int(n_nodes), int(n_rel), trip.shape[0]
        self.rank, self.world = rank, world
        indeg = np.bincount(trip[:, 0], minlength=n_nodes).astype(np.int64)
        if row_ns > 0 and world > 1 and blended:
            both = indeg + np.bincount(trip[:, 2], minlength=n_nodes)
            self.bounds = partition_rows_blended(row_ns + 0.5 * edge_ns * both, world)
        elif row_ns > 0 and world > 1:
            both = indeg + np.bincount(trip[:, 2], minlength=n_nodes)
            self.bounds = partition_rows_minmax(both, world, row_ns, edge_ns / 2)
        else:
            self.bounds = partition_rows(indeg, world)
        self.r0, self.r1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.eid_in = np.flatnonzero((trip[:, 0] >= self.r0) & (trip[:, 0] < self.r1))     # head owned: forward
        self.eid_out = np.flatnonzero((trip[:, 2] >= self.r0) & (trip[:, 2] < self.r1))    # tail owned: backward
        self.g_in = CKGraph(n_nodes, trip[self.eid_in], n_rel=n_rel)
        self.g_out = CKGraph(n_nodes, trip[self.eid_out], n_rel=n_rel)
        # Global CSR order (rank of an edge under (head, rel, eid)) is the concatenation of the ranks'
        # local CSR orders because ranks own contiguous head ranges.  csr_rank[e] = global CSR position.
        order = np.argsort(trip[:, 0].astype(np.int64) * n_rel + trip[:, 1], kind="stable")
        self.csr_rank = np.empty(trip.shape[0], dtype=np.int64)
        self.csr_rank[order] = np.arange(trip.shape[0], dtype=np.int64)
        # edge-count boundaries of the ranks inside the global CSR order
        self.edge_bounds = np.searchsorted(trip[order, 0], self.bounds, side="left").astype(np.int64)
        self.edge_bounds[-1] = trip.shape[0]


def exchange_rows(buf: th.Tensor, bounds, group=None):
    """All-gather of row ranges, in place: rank k's rows ``bounds[k]:bounds[k+1]`` of ``buf`` reach
    everyone.  Ranges are uneven (edge-balanced).  NCCL: every rank pads its rows to the largest range
    and ONE ``all_gather_into_tensor`` moves them (a real ring/tree all-gather instead of one
    broadcast per owner), then a single ``cat`` drops the padding into place.  gloo (CPU tests) gets
    one broadcast per owner."""
    world = len(bounds) - 1
    if tdist.get_backend(group) == "nccl":
        rank = tdist.get_rank(group)
        sizes = [int(bounds[k + 1]) - int(bounds[k]) for k in range(world)]
        m = max(sizes)
        tail = tuple(buf.shape[1:])
        stage_in = th.empty((m,) + tail, dtype=buf.dtype, device=buf.device)
        stage_in[: sizes[rank]] = buf[int(bounds[rank]): int(bounds[rank + 1])]
        stage_out = th.empty((world, m) + tail, dtype=buf.dtype, device=buf.device)
        tdist.all_gather_into_tensor(stage_out, stage_in, group=group)
        th.cat([stage_out[k, : sizes[k]] for k in range(world)], out=buf)
        return buf
    for k in range(world):
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

    def __init__(self, model, n_nodes, triplets, n_rel, device, group=None, exchange=None, max_batch=16384):
        self.model, self.device, self.group = model, th.device(device), group
        self.rank, self.world = tdist.get_rank(group), tdist.get_world_size(group)
        exchange = exchange or os.environ.get("KGAT_DIST_EXCHANGE", "peer")
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl', got %r" % (exchange,))
        self.exchange = exchange
        # cost model of the partition (measured on B200, Amazon-book shape, ns per step): an edge costs ~0.36
        # (all gathers of a step), a row ~1 for its dense work plus, with the peer exchange, ~2.9 per peer for
        # the NVLink copies of its exchanged rows (6 buffers, 1.3 KB per row and peer at ~450 GB/s)
        self.peer = None
        if exchange == "peer":
            dims = [model.entity_embed.weight.shape[1]] + [l.res_fc_2.weight.shape[0] for l in model.layers]
            N, E, L = int(n_nodes), int(np.asarray(triplets).shape[0]), len(dims) - 1
            floats = E + N * sum(dims[1:L]) + N * sum(dims[:L]) + N * dims[0] \
                + self.world * sum(dims[l] * dims[l + 1] for l in range(L)) + 3 * int(max_batch) * sum(dims)
            self.peer = PeerArena(4 * floats + 256 * (2 * L + 8), self.device, group)
            self.peer.max_batch = int(max_batch)          # rows of the loss-row table (ops.DistBprLoss)
        # cost model of the partition, per step (measured on B200s, Amazon-book shape, profiles/README.md): a rank's
        # kernel time is ~3 ns per owned row of dense work + ~3 ns per row and DESTINATION of its exchanged rows
        # (1.3 KB per row: one destination with the NVSwitch multicast, world - 1 with unicast pushes) + 0.29 ns per
        # in-edge, and the step takes as long as its slowest rank
        row_ns = os.environ.get("KGAT_DIST_ROW_NS")
        if row_ns is not None:
            row_ns = float(row_ns)
        elif self.peer is None:
            row_ns = 3.0
        else:
            row_ns = 3.0 + 3.0 * (1 if self.peer.multicast else self.world - 1)
        self.part = Partition(n_nodes, triplets, n_rel, self.rank, self.world, row_ns=row_ns, edge_ns=0.29,
                              blended=os.environ.get("KGAT_DIST_PARTITION", "blended") == "blended")
        if exchange == "peer":
            # fixed allocation order = identical offsets on every rank
            self.peer.buf(("w",), (max(E, 1),))
            for l in range(L - 1):
                self.peer.buf(("h", l), (N, dims[l + 1]))
            for l in range(L):
                self.peer.buf(("dh", l), (N, dims[l]))
            self.peer.buf(("d_entity",), (N, dims[0]))
            self.peer.buf(("dW2",), (self.world, sum(dims[l] * dims[l + 1] for l in range(L))))
        self.dg_in = self.part.g_in.device_graph(self.device)
        self.dg_out = self.part.g_out.device_graph(self.device)
        p = self.part
        # positions (in the global CSR order) of: every edge by edge id, and this rank's out-edges in the
        # order edge_weights() of the local out-graph expects them (its local edge-id order)
        self._csr_rank = th.from_numpy(p.csr_rank).to(self.device)
        self._out_rank = th.from_numpy(p.csr_rank[p.eid_out]).to(self.device)
        # the logits kernel is chosen from the GLOBAL (head, relation)-run count (runs never straddle ranks: a rank
        # owns whole head rows), so every rank and the single-GPU step run the same arithmetic
        self.att_impl = None
        d0 = model.entity_embed.weight.shape[1]
        if d0 == 64 and model.relation_embed.weight.shape[1] == 64 and self.part.n_edges > 0 and n_rel <= 1024 \
                and self.device.type == "cuda":
            self.dg_in.runs()
            counts = th.from_numpy(np.stack([self.dg_in.rel_edges, self.dg_in.rel_runs])).to(self.device)
            tdist.all_reduce(counts, group=group)
            counts = counts.cpu().numpy()
            self.att_impl, mode = ops.attention_plan_for(counts[0], counts[1])
            self.dg_in.set_rel_mode(mode)         # per-relation marks of the whole graph, not of this rank's slice
        self.w_full = None
        self._plan = None
        self._dw_views = None
        self._calls = 0
        self.poll_every = int(os.environ.get("KGAT_PEER_POLL_EVERY", "128"))

    def poll_peers(self, force=False):
        """Every ``poll_every``-th call (or when forced): synchronise and raise if a peer barrier timed out."""
        if self.peer is None or th.cuda.is_current_stream_capturing():
            return
        self._calls += 1
        if force or (self.poll_every > 0 and self._calls % self.poll_every == 0):
            self.peer.check()

    # attention-update mode ---------------------------------------------------------------
    def compute_attention(self):
        """Attention refresh (models.py:146-154) on the partitioned graph; ``attention_weights()``
        returns the (E, 1) tensor in original edge order when the caller wants it."""
        m, p = self.model, self.part
        e0, e1 = int(p.edge_bounds[self.rank]), int(p.edge_bounds[self.rank + 1])
        with th.no_grad():
            if self.peer is not None:
                # the softmax kernel writes this rank's slice of the all-edges vector (global CSR order)
                # into every arena.  The forward pass only needs the weights of the OWN in-edges (local), so no
                # barrier follows here: the peers' slices are first read when the CSC layout of the out-edges
                # is built for the backward pass, and by then the barrier after the first layer's row exchange
                # has already covered these stores (a release is cumulative over the stream) - one barrier less
                # per step.  A single-layer model has no such barrier and gets an explicit one (finish_weights).
                w_glob = self.peer.buf(("w",), (max(p.n_edges, 1),))[: p.n_edges]
                _, w_csr = ops.attention(self.dg_in, m.entity_embed.weight, m.relation_embed.weight, m.W_R,
                                         rows=(p.r0, p.r1), want_eid=False, w_csr_out=w_glob[e0:e1], peer=self.peer,
                                         impl=self.att_impl)
                w_csc = None
            else:
                _, w_csr = ops.attention(self.dg_in, m.entity_embed.weight, m.relation_embed.weight, m.W_R,
                                         rows=(p.r0, p.r1), want_eid=False, impl=self.att_impl)
                # one uneven all-gather assembles the weights of ALL edges in global CSR order
                w_glob = th.empty(p.n_edges, dtype=th.float32, device=self.device)
                w_glob[e0:e1] = w_csr[: self.dg_in.n_edges]
                exchange_rows(w_glob, p.edge_bounds, self.group)
                _, w_csc = self.dg_out.edge_weights(w_glob[self._out_rank])
        self._w_glob = w_glob
        self._w_complete = self.peer is None
        self.w_full = None
        if self.peer is not None:
            self._plan = ops.PropagatePlan(self.dg_in, w_csr, w_csc, dg_bwd=self.dg_out, rows=(p.r0, p.r1),
                                           peer=self.peer, finish_weights=self._finish_weights)
        else:
            self._plan = ops.PropagatePlan(
                self.dg_in, w_csr, w_csc, dg_bwd=self.dg_out, rows=(p.r0, p.r1),
                exchange=lambda t: exchange_rows(t, p.bounds, self.group),
                reduce=self._reduce_weight_grads)
        return self

    def _finish_weights(self, barrier_done):
        """Peer exchange: build the CSC layout of the out-edge weights once every rank's slice has arrived.
        ``barrier_done``: a peer barrier has been enqueued since the softmax (the caller's row exchange)."""
        with th.no_grad():
            if not barrier_done:
                self.peer.barrier()
            self._w_complete = True
            _, w_csc = self.dg_out.edge_weights(self._w_glob[self._out_rank])
        return w_csc

    def attention_weights(self):
        """(E, 1) weights in original edge order (what Model.compute_attention returns)."""
        if not self._w_complete:                      # peer exchange: the peers' slices need a barrier first
            self._plan.w_csc = self._finish_weights(False)
        if self.w_full is None:
            self.w_full = self._w_glob[self._csr_rank].unsqueeze(1)
        return self.w_full

    def _reduce_weight_grads(self, grads):
        """ONE all-reduce for all layers' dW2 (called once per backward with the list of tensors)."""
        flat = th.cat([g.reshape(-1) for g in grads])
        tdist.all_reduce(flat, group=self.group)
        off = 0
        for g in grads:
            g.copy_(flat[off: off + g.numel()].view_as(g))
            off += g.numel()

    # CF-embedding mode -------------------------------------------------------------------
    def gnn(self, drop_masks=None):
        """(N, sum d) embedding whose OWNED rows are valid (the others are zero with the NCCL exchange
        and undefined with the peer exchange; ``full_embedding`` fills them); differentiable."""
        if self._plan is None:
            raise RuntimeError("call compute_attention() first")
        self.poll_peers()
        m = self.model
        weights = [l.res_fc_2.weight for l in m.layers]
        if drop_masks is None:
            # the full-size masks Model.gnn would draw (models.py:70): every rank draws the same ones from the same
            # generator state, so the partitioned step drops the elements the single-GPU step drops
            from .models import _dropout_mask
            drop_masks = [_dropout_mask(l.mess_drop, self.part.n_nodes, l.res_fc_2.out_features, self.device)
                          for l in m.layers]
        return ops.propagate(self._plan, list(drop_masks), False, m.entity_embed.weight, *weights)

    def get_loss(self, embedding, src_ids, pos_dst_ids, neg_dst_ids):
        """BPR loss of models.py:170-178 on rows gathered from their owners; replicated value.  The
        gradient stays compact and local (ops.DistBprLoss)."""
        return ops.DistBprLoss.apply(embedding, src_ids, pos_dst_ids, neg_dst_ids, self.model._reg_lambda_gnn,
                                     getattr(embedding, "_kgat_channel", None), (self.part.r0, self.part.r1),
                                     self.group, self.peer)

    def close(self):
        """Collective teardown of the peer arenas (no-op for the NCCL exchange)."""
        if self.peer is not None:
            self._plan = None
            self.peer.close()

    def full_embedding(self, embedding):
        with th.no_grad():
            return exchange_rows(embedding.detach().clone(), self.part.bounds, self.group)

    # evaluation ----------------------------------------------------------------------------
    def calc_recall_ndcg(self, embedding_full, train_user_dict, test_user_dict, item_id_range, K):
        """metric.calc_recall_ndcg with the eval users sharded over ranks (metric.py:36-69)."""
        self.poll_peers(force=True)
        users = list(test_user_dict.keys())
        mine = users[self.rank::self.world]
        sub = {u: test_user_dict[u] for u in mine}
        sums = th.zeros(3, dtype=th.float64, device=self.device)
        if mine:
            topk, plan = metric.topk_items(embedding_full, train_user_dict, sub, item_id_range, K,
                                           plan=metric.EvalPlan(train_user_dict, sub, len(item_id_range)))
            rec, ndcg = metric.metrics_from_topk_device(topk, plan, K)
            sums += th.tensor([rec * len(mine), ndcg * len(mine), len(mine)], dtype=th.float64, device=self.device)
        tdist.all_reduce(sums, group=self.group)
        n = float(sums[2].item())
        return float(sums[0].item()) / n, float(sums[1].item()) / n
