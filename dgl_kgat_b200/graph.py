"""Collaborative-knowledge-graph object: the drop-in for the DGLGraph the reference builds.

The reference creates ``DGLGraph(); add_nodes(N); add_edges(triplet[:,2], triplet[:,0]);
readonly(); ndata['id'] = arange(N); edata['type'] = triplet[:,1]`` (dataset.py:112-130)
and the model/driver touch only ``ndata['id']``, ``edata['type']``, a settable
``edata['w']`` and ``local_var()`` (models.py:61,148,157; kgat.py:96-109,144).
``CKGraph`` offers exactly that surface; underneath it owns the destination-sorted
CSR, the source-sorted CSC and the relation buckets that ``kgat_graph_build`` produces
on the device (include/kgat_b200.h).  Edge id == triplet row, src = tail, dst = head.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch as th

from . import _lib

__all__ = ["CKGraph", "DeviceGraph"]


class DeviceGraph:
    """The device-resident structure of one graph on one GPU (all int32 tensors)."""

    def __init__(self, n_nodes, n_rel, triplets: np.ndarray, device):
        lib = _lib.load()
        if device.type != "cuda":
            raise _lib.KgatError("the KGAT graph structure lives on a CUDA device; got %s" % device)
        with th.cuda.device(device):
            _lib.check(lib.kgat_device_check(), "kgat_device_check")
            E = int(triplets.shape[0])
            self.n_nodes, self.n_edges, self.n_rel, self.device = int(n_nodes), E, int(n_rel), device
            cols = th.from_numpy(np.ascontiguousarray(triplets.T)).to(device)      # (3, E) int32
            self.arrays = {}
            for name in _lib.GRAPH_ARRAYS:
                n = {"csr_ptr": n_nodes + 1, "csc_ptr": n_nodes + 1, "rel_ptr": n_rel + 1}.get(name, max(E, 1))
                self.arrays[name] = th.empty(n, dtype=th.int32, device=device)
            self.struct = _lib.KgatGraph(self.n_nodes, E, self.n_rel,
                                         *[self.arrays[n].data_ptr() for n in _lib.GRAPH_ARRAYS])
            self.struct.split = self.struct.split_scratch = None
            ws_bytes = lib.kgat_graph_build_workspace_bytes(self.n_nodes, E, self.n_rel)
            ws = th.empty(max(ws_bytes, 16), dtype=th.uint8, device=device)
            _lib.check(lib.kgat_graph_build(ctypes.byref(self.struct), cols[0].data_ptr(), cols[1].data_ptr(),
                                            cols[2].data_ptr(), ws.data_ptr(), ws_bytes, _lib.stream_ptr()),
                       "kgat_graph_build")
            # Rows of more than SPLIT_ROW edges (hot entities of a power-law KG) are cut into segments that the
            # aggregation / softmax kernels hand to one CTA each; the tables are tiny and usually empty.  The call
            # synchronises, which also keeps cols / ws alive until the asynchronous build has read them.
            self.split_table = th.empty(lib.kgat_graph_split_ints(E), dtype=th.int32, device=device)
            self.struct.split = self.split_table.data_ptr()
            counts = (ctypes.c_int32 * 4)()
            _lib.check(lib.kgat_graph_split_build(ctypes.byref(self.struct), counts, _lib.stream_ptr()),
                       "kgat_graph_split_build")
            self.split_counts = tuple(int(c) for c in counts)
            (self.struct.n_csr_seg, self.struct.n_csr_long, self.struct.n_csc_seg,
             self.struct.n_csc_long) = self.split_counts
            n_seg = max(self.struct.n_csr_seg, self.struct.n_csc_seg)
            if n_seg > 0 and os.environ.get("KGAT_SPLIT_ROWS", "1") != "0":
                self.split_scratch = th.empty(n_seg * _lib.SPLIT_SCRATCH_WIDTH, dtype=th.float32, device=device)
                self.struct.split_scratch = self.split_scratch.data_ptr()

    @property
    def ref(self):
        return ctypes.byref(self.struct)

    def runs(self):
        """(head, relation) runs of the relation buckets (``kgat_runs_build``), built on first use."""
        if getattr(self, "_runs", None) is None:
            lib = _lib.load()
            E, R = self.n_edges, self.n_rel
            with th.cuda.device(self.device):
                sizes = {"bk_run": max(E, 1), "run_ptr": E + 1, "run_head": max(E, 1), "rel_run_ptr": R + 1, "n_runs": 1,
                         "rel_mode": R}
                self.run_arrays = {n: th.zeros(sizes[n], dtype=th.int32, device=self.device) for n in _lib.RUN_ARRAYS}
                self._runs = _lib.KgatRuns(*[self.run_arrays[n].data_ptr() for n in _lib.RUN_ARRAYS])
                ws_bytes = lib.kgat_runs_build_workspace_bytes(E)
                ws = th.empty(max(ws_bytes, 16), dtype=th.uint8, device=self.device)
                _lib.check(lib.kgat_runs_build(self.ref, ctypes.byref(self._runs), ws.data_ptr(), ws_bytes,
                                               _lib.stream_ptr()), "kgat_runs_build")
                th.cuda.current_stream().synchronize()      # ws must outlive the build
                self.n_runs = int(self.run_arrays["n_runs"].item())
                # per relation: edges, runs, and the mark the library derived from them (1 = run shortcut pays)
                rel_ptr = self.arrays["rel_ptr"].cpu().numpy().astype(np.int64)
                rel_run_ptr = self.run_arrays["rel_run_ptr"].cpu().numpy().astype(np.int64)
                self.rel_edges, self.rel_runs = np.diff(rel_ptr), np.diff(rel_run_ptr)
                self._rel_ptr, self._rel_run_ptr = rel_ptr, rel_run_ptr
                self.rel_mode = self.run_arrays["rel_mode"].cpu().numpy().copy()
        return ctypes.byref(self._runs)

    def set_rel_mode(self, mode):
        """Override the per-relation marks (dist.DistKGAT: decided from the GLOBAL counts so that every rank and the
        single-GPU step run the same arithmetic on every edge)."""
        self.runs()
        mode = np.ascontiguousarray(mode, dtype=np.int32)
        if mode.shape != (self.n_rel,):
            raise ValueError("rel_mode must have one entry per relation")
        self.run_arrays["rel_mode"].copy_(th.from_numpy(mode))
        self.rel_mode = mode.copy()

    def hybrid_spans(self):
        """What kgat_attention_logits_hybrid needs about the marked relations of THIS graph: the run range and the
        bucket-position range from the first to the last marked relation that has edges here, or None."""
        self.runs()
        marked = np.flatnonzero((self.rel_mode == 1) & (self.rel_edges > 0))
        if len(marked) == 0:
            return None
        lo, hi = int(marked[0]), int(marked[-1])
        return {"v_base": int(self._rel_run_ptr[lo]), "v_count": int(self._rel_run_ptr[hi + 1] - self._rel_run_ptr[lo]),
                "p_begin": int(self._rel_ptr[lo]), "p_end": int(self._rel_ptr[hi + 1]),
                "marked_edges": int(self.rel_edges[marked].sum()), "marked_runs": int(self.rel_runs[marked].sum())}

    def edge_weights(self, w_eid: th.Tensor):
        """(w_csr, w_csc) for an (E,) or (E,1) weight tensor given in edge-id order."""
        lib = _lib.load()
        w = w_eid.detach().reshape(-1).to(device=self.device, dtype=th.float32).contiguous()
        if w.numel() != self.n_edges:
            raise ValueError("edge weight tensor has %d rows, graph has %d edges" % (w.numel(), self.n_edges))
        w_csr = th.empty(max(self.n_edges, 1), dtype=th.float32, device=self.device)
        w_csc = th.empty(max(self.n_edges, 1), dtype=th.float32, device=self.device)
        with th.cuda.device(self.device):
            _lib.check(lib.kgat_edge_weights_from_eid(self.ref, w.data_ptr(), w_csr.data_ptr(), w_csc.data_ptr(),
                                                      _lib.stream_ptr()), "kgat_edge_weights_from_eid")
        return w_csr, w_csc


class _Core:
    """State shared by a graph and its ``local_var()`` views."""

    def __init__(self, n_nodes, triplets, n_rel):
        trip = np.ascontiguousarray(np.asarray(triplets), dtype=np.int32)
        if trip.ndim != 2 or trip.shape[1] != 3:
            raise ValueError("triplets must be an (E, 3) array of (head, relation, tail)")
        self.n_nodes = int(n_nodes)
        E = trip.shape[0]
        if E:
            if trip[:, [0, 2]].min() < 0 or trip[:, [0, 2]].max() >= self.n_nodes:
                raise ValueError("triplet entity id outside [0, n_nodes)")
            if trip[:, 1].min() < 0:
                raise ValueError("negative relation id")
        self.n_rel = int(n_rel) if n_rel is not None else (int(trip[:, 1].max()) + 1 if E else 1)
        if E and trip[:, 1].max() >= self.n_rel:
            raise ValueError("relation id outside [0, n_rel)")
        self.triplets = trip
        self.n_edges = E
        self._dev = {}
        self._w_cache = None       # (key, w_eid kept alive so its address cannot be recycled, w_csr, w_csc)

    def device_graph(self, device) -> DeviceGraph:
        device = th.device(device)
        if device.type == "cuda" and device.index is None:
            device = th.device("cuda", th.cuda.current_device())
        dg = self._dev.get(device)
        if dg is None:
            dg = DeviceGraph(self.n_nodes, self.n_rel, self.triplets, device)
            self._dev[device] = dg
        return dg

    @staticmethod
    def _key(w, device):
        return (w.data_ptr(), w._version, tuple(w.shape), str(device))

    def remember_weights(self, w_eid, device, w_csr, w_csc):
        self._w_cache = (self._key(w_eid, self.device_graph(device).device), w_eid, w_csr, w_csc)

    def weights_for(self, w_eid, device):
        dg = self.device_graph(device)
        key = self._key(w_eid, dg.device)
        if self._w_cache is not None and self._w_cache[0] == key:
            return self._w_cache[2], self._w_cache[3]
        w_csr, w_csc = dg.edge_weights(w_eid)
        self._w_cache = (key, w_eid, w_csr, w_csc)
        return w_csr, w_csc


class CKGraph:
    def __init__(self, n_nodes, triplets, n_rel=None, _core=None):
        self._core = _core if _core is not None else _Core(n_nodes, triplets, n_rel)
        self.ndata, self.edata = {}, {}
        if _core is None:
            self.ndata["id"] = th.arange(self._core.n_nodes, dtype=th.long)            # dataset.py:118
            self.edata["type"] = th.from_numpy(self._core.triplets[:, 1].astype(np.int64))  # dataset.py:119

    # --- the DGLGraph surface the reference uses -------------------------------------------
    def local_var(self):
        g = CKGraph(None, None, _core=self._core)
        g.ndata, g.edata = dict(self.ndata), dict(self.edata)
        return g

    def number_of_nodes(self):
        return self._core.n_nodes

    def number_of_edges(self):
        return self._core.n_edges

    def readonly(self, state=True):          # graphs are always immutable here
        return self

    def all_edges(self, order="eid"):
        t = self._core.triplets
        return th.from_numpy(t[:, 2].astype(np.int64)), th.from_numpy(t[:, 0].astype(np.int64))

    # --- structure access for the kernels ---------------------------------------------------
    @property
    def n_nodes(self):
        return self._core.n_nodes

    @property
    def n_edges(self):
        return self._core.n_edges

    @property
    def n_rel(self):
        return self._core.n_rel

    @property
    def triplets(self):
        return self._core.triplets

    def device_graph(self, device) -> DeviceGraph:
        return self._core.device_graph(device)

    def edge_weight_layouts(self, device):
        """(w_csr, w_csc) of the weights currently stored as ``edata['w']``."""
        if "w" not in self.edata:
            raise KeyError("edata['w'] is not set: call compute_attention(g) and store the result "
                           "with g.edata['w'] = A_w first (kgat.py:143-144)")
        return self._core.weights_for(self.edata["w"], device)

    def ids_are_identity(self):
        ids = self.ndata["id"]
        cached = getattr(self._core, "_identity_check", None)
        key = (ids.data_ptr(), ids._version)
        if cached is not None and cached[0] == key:
            return cached[1]
        ok = ids.numel() == self.n_nodes and bool(
            th.equal(ids.reshape(-1).cpu(), th.arange(self.n_nodes, dtype=ids.dtype)))
        self._core._identity_check = (key, ok)
        return ok

    def __repr__(self):
        return "CKGraph(num_nodes=%d, num_edges=%d, num_rel=%d)" % (self.n_nodes, self.n_edges, self.n_rel)
