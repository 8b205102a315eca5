"""numpy restatement of the graph structures the CUDA path builds (bit-exact target).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).

The reference hands the (E, 3) ``(h, r, t)`` triplet array to DGL as
``add_edges(src=t, dst=h)`` with ``edata['type'] = r`` (dataset.py:114-119) and
lets DGL derive its CSR internally; edge id == triplet row.  The B200 path makes
that structure explicit.  Every array below is a pure function of the triplet
array, defined with stable sorts so there is exactly one right answer:

CSR  (destination-sorted)   position c = rank of edge under key (head, rel, eid)
     csr_ptr[N+1]  csr_src[c]=tail  csr_dst[c]=head  csr_rel[c]  csr_eid[c]=triplet row
CSC  (source-sorted, for the transposed aggregation in backward)
     position k = rank of CSR position under key (tail, c)
     csc_ptr[N+1]  csc_dst[k]=head  csc_pos[k]=c
relation buckets (for the relation-grouped projection)
     position p = rank of CSR position under key (rel, c)
     rel_ptr[R+1]  bk_pos[p]=c  bk_head[p]  bk_tail[p]
     Bucket r holds exactly the edges ``filter_edges(type == r)`` yields
     (models.py:150); inside a bucket they are ordered by head so that one
     tile of a bucket re-uses head rows.
"""
from __future__ import annotations

import numpy as np


def build_structure(trip: np.ndarray, n_nodes: int, n_rel: int) -> dict:
    trip = np.asarray(trip)
    h = trip[:, 0].astype(np.int64)
    r = trip[:, 1].astype(np.int64)
    t = trip[:, 2].astype(np.int64)
    E = trip.shape[0]
    order = np.argsort(h * n_rel + r, kind="stable")
    csr_dst = h[order]
    csr_src = t[order]
    csr_rel = r[order]
    csr_ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(h, minlength=n_nodes), out=csr_ptr[1:])

    corder = np.argsort(csr_src, kind="stable")
    csc_ptr = np.zeros(n_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(t, minlength=n_nodes), out=csc_ptr[1:])

    border = np.argsort(csr_rel, kind="stable")
    rel_ptr = np.zeros(n_rel + 1, dtype=np.int64)
    np.cumsum(np.bincount(r, minlength=n_rel), out=rel_ptr[1:])

    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return {
        "csr_ptr": i32(csr_ptr), "csr_src": i32(csr_src), "csr_dst": i32(csr_dst),
        "csr_rel": i32(csr_rel), "csr_eid": i32(order),
        "csc_ptr": i32(csc_ptr), "csc_dst": i32(csr_dst[corder]), "csc_pos": i32(corder),
        "rel_ptr": i32(rel_ptr), "bk_pos": i32(border),
        "bk_head": i32(csr_dst[border]), "bk_tail": i32(csr_src[border]),
        "n_edges": E,
    }


def partition_rows(csr_ptr: np.ndarray, n_parts: int) -> np.ndarray:
    """Contiguous destination ranges with balanced edge counts: boundary b_k is the
    first row whose prefix edge count reaches k * E / n_parts."""
    E = int(csr_ptr[-1])
    targets = (np.arange(n_parts + 1, dtype=np.int64) * E) // n_parts
    b = np.searchsorted(csr_ptr, targets, side="left").astype(np.int64)
    b[0], b[-1] = 0, len(csr_ptr) - 1
    return np.maximum.accumulate(b)


def build_runs(struct, n_rel):
    """(head, relation) runs of the relation buckets: maximal groups of consecutive bucket positions with the
    same relation and head (the bucket order is relation, head, edge id).  Restates what
    kgat_runs_build produces: bk_run, run_ptr (n_runs + 1 entries), run_head, rel_run_ptr, n_runs."""
    import numpy as np
    bk_pos, bk_head, rel_ptr = struct["bk_pos"], struct["bk_head"], struct["rel_ptr"]
    E = len(bk_pos)
    rel = struct["csr_rel"][bk_pos] if E else np.zeros(0, np.int32)
    flags = np.ones(E, dtype=np.int32)
    if E > 1:
        flags[1:] = ((bk_head[1:] != bk_head[:-1]) | (rel[1:] != rel[:-1])).astype(np.int32)
    bk_run = (np.cumsum(flags) - 1).astype(np.int32)
    n_runs = int(flags.sum())
    starts = np.flatnonzero(flags).astype(np.int32)
    run_ptr = np.append(starts, E).astype(np.int32)
    run_head = bk_head[starts].astype(np.int32) if E else np.zeros(0, np.int32)
    rel_run_ptr = np.array([bk_run[p] if p < E else n_runs for p in rel_ptr[: n_rel + 1]], dtype=np.int32)
    return {"bk_run": bk_run, "run_ptr": run_ptr, "run_head": run_head, "rel_run_ptr": rel_run_ptr, "n_runs": n_runs}


def split_table(ptr, split_row, split_seg):
    """Segment table of one side of the graph (kgat_graph_split_build, include/kgat_b200.h): every row with more
    than ``split_row`` edges, ascending, is cut into n = ceil(deg / split_seg) segments of ceil(deg / n) edges
    rounded up to a multiple of 32, in edge order.  This structure is not in the reference (DGL's kernels walk
    a row with one thread block at most); it only decides which CTA sums which edges."""
    ptr = np.asarray(ptr, dtype=np.int64)
    deg = np.diff(ptr)
    rows = np.flatnonzero(deg > split_row)
    out = {k: [] for k in ("seg_row", "seg_beg", "seg_end", "seg_long", "long_row", "long_seg0", "long_nseg")}
    for li, r in enumerate(rows):
        d, beg = int(deg[r]), int(ptr[r])
        n = -(-d // split_seg)
        per = (-(-d // n) + 31) & ~31
        out["long_row"].append(int(r))
        out["long_seg0"].append(len(out["seg_row"]))
        out["long_nseg"].append(n)
        for i in range(n):
            b = min(beg + d, beg + i * per)
            out["seg_row"].append(int(r))
            out["seg_beg"].append(b)
            out["seg_end"].append(min(beg + d, b + per))
            out["seg_long"].append(li)
    return {k: np.asarray(v, dtype=np.int32) for k, v in out.items()}
