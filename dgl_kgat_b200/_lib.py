"""ctypes binding of libkgat_b200.so (the C ABI declared in include/kgat_b200.h).

There is no CPU fallback: if the shared library is missing or the current device
is not an sm_100 GPU every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libkgat_b200.so")

c_i32p = C.c_void_p     # device pointers travel as plain addresses
c_f32p = C.c_void_p


class KgatGraph(C.Structure):
    """Mirror of ``kgat_graph_t``."""
    _fields_ = [("n_nodes", C.c_int64), ("n_edges", C.c_int64), ("n_rel", C.c_int64)] + [
        (name, C.c_void_p) for name in (
            "csr_ptr", "csr_src", "csr_dst", "csr_rel", "csr_eid",
            "csc_ptr", "csc_dst", "csc_pos", "csc_eid",
            "rel_ptr", "bk_pos", "bk_head", "bk_tail")] + [
        # row splitting (kgat_graph_split_build): segment tables, partial-sum scratch, host copies of the sizes
        ("split", C.c_void_p), ("split_scratch", C.c_void_p),
        ("n_csr_seg", C.c_int32), ("n_csr_long", C.c_int32), ("n_csc_seg", C.c_int32), ("n_csc_long", C.c_int32)]


GRAPH_ARRAYS = [name for name, _ in KgatGraph._fields_[3:16]]
SPLIT_ROW, SPLIT_SEG, SPLIT_SCRATCH_WIDTH = 256, 512, 128     # KGAT_SPLIT_* of the header


class KgatRuns(C.Structure):
    """Mirror of ``kgat_runs_t``."""
    _fields_ = [(name, C.c_void_p) for name in ("bk_run", "run_ptr", "run_head", "rel_run_ptr", "n_runs", "rel_mode")]


RUN_ARRAYS = [name for name, _ in KgatRuns._fields_]

MAX_RANKS = 8                   # KGAT_MAX_RANKS
PEER_HEADER_BYTES = 1024        # KGAT_PEER_HEADER_BYTES
IPC_HANDLE_BYTES = 64           # KGAT_IPC_HANDLE_BYTES


class KgatPeerGroup(C.Structure):
    """Mirror of ``kgat_peer_group_t``."""
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("arena_bytes", C.c_uint64),
                ("arena", C.c_void_p * MAX_RANKS), ("multicast", C.c_void_p)]


# name -> (restype, argtypes); must list every KGAT_API symbol of the header
_GP = C.POINTER(KgatGraph)
_PG = C.POINTER(KgatPeerGroup)
_RP = C.POINTER(KgatRuns)
SIGNATURES = {
    "kgat_version": (C.c_int, []),
    "kgat_error_string": (C.c_char_p, [C.c_int]),
    "kgat_last_cuda_error": (C.c_int, []),
    "kgat_device_check": (C.c_int, []),
    "kgat_graph_build_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int64]),
    "kgat_graph_build": (C.c_int, [_GP, c_i32p, c_i32p, c_i32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kgat_graph_split_ints": (C.c_size_t, [C.c_int64]),
    "kgat_graph_split_build": (C.c_int, [_GP, C.POINTER(C.c_int32), C.c_void_p]),
    "kgat_attention_logits": (C.c_int, [_GP, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, c_f32p, C.c_void_p]),
    "kgat_attention_fallbacks": (C.c_int, []),
    "kgat_runs_build_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "kgat_runs_build": (C.c_int, [_GP, _RP, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kgat_attention_logits_runs": (C.c_int, [_GP, _RP, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, c_f32p, c_f32p,
                                             C.c_void_p]),
    "kgat_attention_logits_hybrid": (C.c_int, [_GP, _RP, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, c_f32p, c_f32p,
                                               C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p]),
    "kgat_segment_softmax": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "kgat_edge_weights_from_eid": (C.c_int, [_GP, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "kgat_layer_fwd": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int,
                                 C.c_float, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int64, C.c_int64,
                                 C.c_void_p]),
    "kgat_layer_bwd_partials": (C.c_int, []),
    "kgat_layer_bwd_reduce": (C.c_int, [C.c_int64, C.c_int, C.c_int, c_f32p, c_f32p, C.c_void_p]),
    "kgat_layer_bwd": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_i32p, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p,
                                 c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_float,
                                 c_f32p, c_f32p, c_f32p, c_f32p, C.c_void_p]),
    "kgat_input_bwd": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_i32p, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p,
                                 C.c_int, c_f32p, C.c_void_p]),
    "kgat_csc_gather": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p, C.c_int, c_f32p, C.c_void_p]),
    "kgat_bpr_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "kgat_bpr_loss_fwd": (C.c_int, [c_f32p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                    c_f32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kgat_bpr_loss_bwd": (C.c_int, [c_f32p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float,
                                    c_f32p, c_f32p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                    c_f32p, c_i32p, c_i32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kgat_transr_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64]),
    "kgat_transr_fwd": (C.c_int, [c_f32p, C.c_int64, c_f32p, c_f32p, C.c_int64, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_float, c_f32p, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "kgat_transr_bwd": (C.c_int, [c_f32p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int64, c_f32p, c_f32p, c_f32p,
                                  C.c_void_p, C.c_size_t, C.c_void_p]),
    "kgat_adam_step": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                 C.c_double, C.c_double, C.c_double, C.c_double, C.c_int64, C.c_void_p]),
    "kgat_peer_arena_create": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "kgat_peer_arena_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "kgat_peer_arena_close": (C.c_int, [C.c_void_p]),
    "kgat_peer_arena_destroy": (C.c_int, [C.c_void_p]),
    "kgat_peer_barrier": (C.c_int, [_PG, C.c_int, C.c_void_p]),
    "kgat_peer_error": (C.c_int, [_PG]),
    "kgat_segment_softmax_push": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p, _PG, C.c_void_p]),
    "kgat_layer_fwd_push": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int,
                                      C.c_float, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int64, C.c_int64, C.c_int64,
                                      _PG, C.c_void_p]),
    "kgat_layer_bwd_push": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_i32p, C.c_int64, C.c_int64, c_f32p, c_f32p,
                                      c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, c_f32p, C.c_int, C.c_int,
                                      C.c_float, c_f32p, c_f32p, c_f32p, c_f32p, _PG, C.c_void_p]),
    "kgat_input_bwd_push": (C.c_int, [_GP, C.c_int64, C.c_int64, c_f32p, c_i32p, C.c_int64, C.c_int64, c_f32p, c_f32p,
                                      c_f32p, C.c_int, c_f32p, _PG, C.c_void_p]),
    "kgat_peer_gather_rows_push": (C.c_int, [c_f32p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int64, C.c_int64,
                                             c_f32p, _PG, C.c_void_p]),
    "kgat_peer_push_slot": (C.c_int, [c_f32p, C.c_int64, c_f32p, _PG, C.c_void_p]),
    "kgat_peer_push_buffer": (C.c_int, [c_f32p, C.c_int64, c_f32p, _PG, C.c_void_p]),
    "kgat_peer_sum_slots": (C.c_int, [c_f32p, C.c_int64, C.c_int, c_f32p, C.c_void_p]),
    "kgat_umma_error_flag": (C.c_int, []),
    "kgat_selftest_umma": (C.c_int, [c_f32p, c_f32p, c_f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_void_p]),
    "kgat_score_topk": (C.c_int, [c_f32p, C.c_int64, C.c_int, c_i32p, C.c_int64, c_i32p, C.c_int64, c_i32p, c_i32p,
                                  C.c_int, c_i32p, c_f32p, C.c_void_p]),
    "kgat_recall_ndcg_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "kgat_recall_ndcg": (C.c_int, [c_i32p, C.c_int64, C.c_int, c_i32p, c_i32p, c_i32p, C.c_void_p, C.c_void_p, C.c_size_t,
                                   C.c_void_p]),
    "kgat_score_topk_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "kgat_score_topk_ws": (C.c_int, [c_f32p, C.c_int64, C.c_int, c_i32p, C.c_int64, c_i32p, C.c_int64, c_i32p, c_i32p,
                                     C.c_int, c_i32p, c_f32p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "kgat_eval_fallbacks": (C.c_int, []),
}


class KgatError(RuntimeError):
    pass


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KgatError(
            "libkgat_b200.so is missing (%s).  Build it with `python -m dgl_kgat_b200.build`; "
            "there is no CPU or PyTorch fallback for the KGAT hot path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        lib = load()
        msg = lib.kgat_error_string(status).decode()
        extra = ""
        if status == -4:
            extra = " (cudaError %d)" % lib.kgat_last_cuda_error()
        raise KgatError("%s failed: %s%s" % (what or "kgat call", msg, extra))


def ptr(t):
    """Device address of a tensor, or None for a null argument."""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


class Recorder:
    """Counts the kernels this library launches and, when ``time_kernels`` is set, brackets every
    C-ABI call with CUDA events on the launching stream (bench.py reads both).  Off by default."""

    def __init__(self):
        self.enabled = False
        self.time_kernels = False
        self.launches = 0
        self.events = {}        # name -> list of (start, stop)

    def reset(self, enabled=True, time_kernels=False):
        self.enabled, self.time_kernels, self.launches, self.events = enabled, time_kernels, 0, {}

    def durations_ms(self):
        """name -> list of per-call durations (call after a synchronize)."""
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.events.items()}


RECORDER = Recorder()


class record:
    """``with record("kgat_layer_fwd", 1): lib.kgat_layer_fwd(...)``"""
    __slots__ = ("name", "n", "start")

    def __init__(self, name, n_launches=1):
        self.name, self.n, self.start = name, n_launches, None

    def __enter__(self):
        r = RECORDER
        if r.enabled:
            r.launches += self.n
            if r.time_kernels:
                import torch
                self.start = torch.cuda.Event(enable_timing=True)
                self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None:
            import torch
            stop = torch.cuda.Event(enable_timing=True)
            stop.record()
            RECORDER.events.setdefault(self.name, []).append((self.start, stop))
        return False
