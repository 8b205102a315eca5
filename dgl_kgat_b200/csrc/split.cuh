// Segment tables of the rows that are split across CTAs (include/kgat_b200.h, kgat_graph_t::split).
//
// Layout of g->split (int32): header of 8 ints {n_csr_seg, n_csr_long, n_csc_seg, n_csc_long, ...}, then per
// side (0 = CSR / in-edges, 1 = CSC / out-edges) seven arrays:
//   seg_row[capS] seg_beg[capS] seg_end[capS] seg_long[capS]   one entry per segment (edge positions [beg, end))
//   long_row[capL] long_seg0[capL] long_nseg[capL]             one entry per split row
// capL = E / KGAT_SPLIT_ROW + 1 (rows above the threshold), capS = E / KGAT_SPLIT_SEG + capL + 1.
#pragma once

#include "common.cuh"

namespace kgat {

constexpr int SPLIT_ROW = KGAT_SPLIT_ROW;
constexpr int SPLIT_SEG = KGAT_SPLIT_SEG;
constexpr int SPLIT_HEADER = 8;

inline int64_t split_cap_long(int64_t E) { return E / SPLIT_ROW + 1; }
inline int64_t split_cap_seg(int64_t E) { return E / SPLIT_SEG + split_cap_long(E) + 1; }
inline int64_t split_side_ints(int64_t E) { return 4 * split_cap_seg(E) + 3 * split_cap_long(E); }

struct SplitSide {
    int32_t *seg_row, *seg_beg, *seg_end, *seg_long, *long_row, *long_seg0, *long_nseg;
    int n_seg, n_long;        // host copies (0 when the graph has no table)
};

inline SplitSide split_side(const kgat_graph_t* g, int side) {
    SplitSide s{};
    if (!g->split) return s;
    const int64_t E = g->n_edges, cS = split_cap_seg(E), cL = split_cap_long(E);
    int32_t* p = g->split + SPLIT_HEADER + side * split_side_ints(E);
    s.seg_row = p; s.seg_beg = p + cS; s.seg_end = p + 2 * cS; s.seg_long = p + 3 * cS;
    p += 4 * cS;
    s.long_row = p; s.long_seg0 = p + cL; s.long_nseg = p + 2 * cL;
    s.n_seg = side == 0 ? g->n_csr_seg : g->n_csc_seg;
    s.n_long = side == 0 ? g->n_csr_long : g->n_csc_long;
    return s;
}

// the pre-pass runs only when the tables say so AND the scratch exists
inline bool split_active(const kgat_graph_t* g, int side) {
    return g->split && g->split_scratch && (side == 0 ? g->n_csr_seg : g->n_csc_seg) > 0;
}

}  // namespace kgat
