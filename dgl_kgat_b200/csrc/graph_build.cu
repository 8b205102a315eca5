// (1) Graph construction: triplets -> destination-sorted CSR, source-sorted CSC and
// relation buckets, all on the device.  Replaces DGLGraph()/add_edges/readonly
// (reference dataset.py:112-130) and the per-relation filter_edges scans
// (reference models.py:150).  Bit-exact against oracle/graph_ref.py: every order is a
// stable LSD radix sort (radix_sort_pairs below) of an explicit key, so the result does not
// depend on scheduling.
#include "common.cuh"
#include "split.cuh"

namespace kgat {

thread_local int g_last_cuda_error = 0;

namespace {

constexpr int TPB = 256;

inline int blocks_for(int64_t n) { return static_cast<int>((n + TPB - 1) / TPB); }

inline int bits_for(uint64_t max_key) {
    int b = 1;
    while (b < 64 && (max_key >> b) != 0) ++b;
    return b;
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

template <typename KeyT>
__global__ void make_edge_keys(const int32_t* __restrict__ head, const int32_t* __restrict__ rel,
                               int64_t n_rel, int64_t n_edges, KeyT* __restrict__ keys,
                               int32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n_edges) {
        keys[i] = static_cast<KeyT>(static_cast<uint64_t>(head[i]) * static_cast<uint64_t>(n_rel) +
                                    static_cast<uint64_t>(rel[i]));
        vals[i] = static_cast<int32_t>(i);
    }
}

__global__ void gather_csr(const int32_t* __restrict__ head, const int32_t* __restrict__ rel,
                           const int32_t* __restrict__ tail, const int32_t* __restrict__ csr_eid,
                           int64_t n_edges, int32_t* __restrict__ csr_src, int32_t* __restrict__ csr_dst,
                           int32_t* __restrict__ csr_rel) {
    int64_t c = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (c < n_edges) {
        int32_t e = csr_eid[c];
        csr_src[c] = tail[e];
        csr_dst[c] = head[e];
        csr_rel[c] = rel[e];
    }
}

__global__ void copy_keys_iota(const int32_t* __restrict__ src, int64_t n, uint32_t* __restrict__ keys,
                               int32_t* __restrict__ vals) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) {
        keys[i] = static_cast<uint32_t>(src[i]);
        vals[i] = static_cast<int32_t>(i);
    }
}

// ptr[v] = number of sorted keys < v, v = 0..n_bins  (segment offsets of a sorted key array)
template <typename KeyT>
__global__ void segment_offsets(const KeyT* __restrict__ sorted, int64_t n, int64_t n_bins,
                                int32_t* __restrict__ ptr) {
    int64_t v = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (v > n_bins) return;
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (static_cast<int64_t>(sorted[mid]) < v) lo = mid + 1; else hi = mid;
    }
    ptr[v] = static_cast<int32_t>(lo);
}

__global__ void gather2(const int32_t* __restrict__ pos, int64_t n, const int32_t* __restrict__ a,
                        const int32_t* __restrict__ b, int32_t* __restrict__ out_a,
                        int32_t* __restrict__ out_b) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) {
        int32_t p = pos[i];
        out_a[i] = a[p];
        out_b[i] = b[p];
    }
}

// ------------------------------------------------------------------------------------------------
// Stable LSD radix sort of (key, value) pairs, 8 bits per pass.
//   pass:  radix_hist   block b counts the digits of its tile of TILE consecutive elements -> hist[digit][b]
//          radix_scan   exclusive scan of hist in (digit, block) order: where block b's run of a digit starts
//          radix_scatter  block b walks its tile in element order, 256 elements per round; the rank of an
//                       element among the equal digits of its round comes from __match_any_sync inside the warp
//                       plus the counts of the lower warps, rounds and blocks before it come from running bases
//                       - so equal keys keep their input order (stable), whatever the scheduling.
// One-off per graph; three passes for the (head, relation) key of the Amazon-book shape.
// ------------------------------------------------------------------------------------------------
constexpr int RADIX_BITS = 8, RADIX = 1 << RADIX_BITS;
constexpr int SORT_THREADS = 256, SORT_ROUNDS = 16, SORT_TILE = SORT_THREADS * SORT_ROUNDS;

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
radix_hist_kernel(const KeyT* __restrict__ keys, int64_t n, int shift, int n_blocks, int32_t* __restrict__ hist) {
    __shared__ int s_h[RADIX];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SORT_TILE;
    for (int r = 0; r < SORT_ROUNDS; ++r) {
        const int64_t i = base + r * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&s_h[static_cast<int>((keys[i] >> shift) & (RADIX - 1))], 1);      // integer counts
    }
    __syncthreads();
    hist[static_cast<size_t>(threadIdx.x) * n_blocks + blockIdx.x] = s_h[threadIdx.x];
}

// one CTA: in-place exclusive scan of the RADIX * n_blocks counters (digit-major), four consecutive counters per
// thread and iteration (4096 per trip through the block-wide scan)
__global__ void __launch_bounds__(1024) radix_scan_kernel(int32_t* __restrict__ hist, int64_t n) {
    __shared__ int s_w[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t c0 = 0; c0 < n; c0 += 4096) {
        const int64_t c = c0 + 4 * static_cast<int64_t>(threadIdx.x);
        int v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = c + j < n ? hist[c + j] : 0;
        const int t = v[0] + v[1] + v[2] + v[3];
        int incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if (lane >= o) incl += u;
        }
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int w = s_w[lane];
            int wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL_MASK, wi, o);
                if (lane >= o) wi += u;
            }
            s_w[lane] = wi - w;
        }
        __syncthreads();
        int run = s_base + s_w[warp] + incl - t;             // exclusive prefix of this thread's first counter
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c + j < n) hist[c + j] = run;
            run += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_base = run;
        __syncthreads();
    }
}

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS)
radix_scatter_kernel(const KeyT* __restrict__ keys_in, const int32_t* __restrict__ vals_in, int64_t n, int shift,
                     int n_blocks, const int32_t* __restrict__ offsets, KeyT* __restrict__ keys_out,
                     int32_t* __restrict__ vals_out) {
    constexpr int WARPS_ = SORT_THREADS / 32;
    __shared__ int s_base[RADIX];                 // where this block's next element of a digit goes
    __shared__ int s_cnt[WARPS_][RADIX];          // per-warp digit counts of the current round
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s_base[threadIdx.x] = offsets[static_cast<size_t>(threadIdx.x) * n_blocks + blockIdx.x];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SORT_TILE;
    for (int r = 0; r < SORT_ROUNDS; ++r) {
#pragma unroll
        for (int w = 0; w < WARPS_; ++w) s_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = base + r * SORT_THREADS + threadIdx.x;
        const bool ok = i < n;
        KeyT k = 0;
        int d = RADIX + lane;                     // out-of-range lanes: a group of their own, never counted
        if (ok) { k = keys_in[i]; d = static_cast<int>((k >> shift) & (RADIX - 1)); }
        const unsigned same = __match_any_sync(FULL_MASK, d);
        const int rank = __popc(same & ((1u << lane) - 1u));
        if (ok && rank == 0) s_cnt[warp][d] = __popc(same);
        __syncthreads();
        if (ok) {
            int before = 0;
            for (int w = 0; w < warp; ++w) before += s_cnt[w][d];
            const int dst = s_base[d] + before + rank;
            keys_out[dst] = k;
            vals_out[dst] = vals_in[i];
        }
        __syncthreads();
        int tot = 0;
#pragma unroll
        for (int w = 0; w < WARPS_; ++w) tot += s_cnt[w][threadIdx.x];
        s_base[threadIdx.x] += tot;
        __syncthreads();
    }
}

struct Workspace {
    size_t keys_a, keys_b, keys_c, vals_a, vals_b, hist, total;
};

inline int sort_blocks(int64_t n) { return static_cast<int>((n + SORT_TILE - 1) / SORT_TILE); }

// Sorts n pairs by the low `bits` bits of the key (stable).  (ka, va) hold the input and are scratch afterwards,
// (kb, vb) is the second pair of buffers; the result lands in (keys_out, vals_out), distinct from all four.
template <typename KeyT>
int radix_sort_pairs(KeyT* ka, KeyT* kb, int32_t* va, int32_t* vb, KeyT* keys_out, int32_t* vals_out, int64_t n,
                     int bits, int32_t* hist, cudaStream_t st) {
    const int nb = sort_blocks(n);
    const int passes = (bits + RADIX_BITS - 1) / RADIX_BITS;
    KeyT *kin = ka, *kout = kb;
    int32_t *vin = va, *vout = vb;
    for (int p = 0; p < passes; ++p) {
        const bool last = p + 1 == passes;
        KeyT* ko = last ? keys_out : kout;
        int32_t* vo = last ? vals_out : vout;
        // the last pass writes straight into the caller's arrays; they must not be the pass's own input
        if (last && (reinterpret_cast<void*>(ko) == reinterpret_cast<void*>(kin) ||
                     reinterpret_cast<void*>(vo) == reinterpret_cast<void*>(vin)))
            return KGAT_ERR_INVALID_ARG;
        radix_hist_kernel<KeyT><<<nb, SORT_THREADS, 0, st>>>(kin, n, p * RADIX_BITS, nb, hist);
        KGAT_LAUNCH_CHECK();
        radix_scan_kernel<<<1, 1024, 0, st>>>(hist, static_cast<int64_t>(RADIX) * nb);
        KGAT_LAUNCH_CHECK();
        radix_scatter_kernel<KeyT><<<nb, SORT_THREADS, 0, st>>>(kin, vin, n, p * RADIX_BITS, nb, hist, ko, vo);
        KGAT_LAUNCH_CHECK();
        kin = ko; vin = vo;
        kout = (kin == ka) ? kb : ka;
        vout = (vin == va) ? vb : va;
    }
    return KGAT_OK;
}

Workspace plan(int64_t n_edges) {
    Workspace w{};
    size_t e = static_cast<size_t>(n_edges > 0 ? n_edges : 1);
    size_t off = 0;
    w.keys_a = off; off += align_up(e * sizeof(uint64_t));
    w.keys_b = off; off += align_up(e * sizeof(uint64_t));
    w.vals_a = off; off += align_up(e * sizeof(int32_t));
    w.vals_b = off; off += align_up(e * sizeof(int32_t));
    w.hist = off; off += align_up(static_cast<size_t>(RADIX) * sort_blocks(static_cast<int64_t>(e)) * sizeof(int32_t));
    // the sorted keys land in a third array: the last pass of a sort must not write into its own input
    w.keys_c = off; off += align_up(e * sizeof(uint64_t));
    w.total = off;
    return w;
}

template <typename KeyT>
int sort_edges_by_head_rel(const kgat_graph_t* g, const int32_t* head, const int32_t* rel, char* ws,
                           const Workspace& w, cudaStream_t st) {
    const int64_t E = g->n_edges;
    KeyT* ka = reinterpret_cast<KeyT*>(ws + w.keys_a);
    KeyT* kb = reinterpret_cast<KeyT*>(ws + w.keys_b);
    int32_t* va = reinterpret_cast<int32_t*>(ws + w.vals_a);
    make_edge_keys<KeyT><<<blocks_for(E), TPB, 0, st>>>(head, rel, g->n_rel, E, ka, va);
    KGAT_LAUNCH_CHECK();
    uint64_t max_key = static_cast<uint64_t>(g->n_nodes) * static_cast<uint64_t>(g->n_rel);
    return radix_sort_pairs<KeyT>(ka, kb, va, reinterpret_cast<int32_t*>(ws + w.vals_b),
                                  reinterpret_cast<KeyT*>(ws + w.keys_c), g->csr_eid, E, bits_for(max_key),
                                  reinterpret_cast<int32_t*>(ws + w.hist), st);
}

// Segment table of one side (split.cuh): one CTA walks the row pointer in chunks of 1024 rows; a block-wide
// exclusive scan of (segments, split rows) per chunk gives every split row its slots, so rows appear in
// ascending order and the segments of a row in edge order - the table is a pure function of the pointer.
// A row of deg > SPLIT_ROW edges gets n = ceil(deg / SPLIT_SEG) segments of ceil(deg / n) edges rounded up
// to a multiple of 32 (the per-warp slices of gather_row_cta stay 32-aligned).
__global__ void __launch_bounds__(1024)
split_table_kernel(const int32_t* __restrict__ ptr, int n_rows, SplitSide sp, int32_t* __restrict__ counts) {
    __shared__ int s_ws[32], s_wl[32];
    __shared__ int s_base_s, s_base_l;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_base_s = 0; s_base_l = 0; }
    __syncthreads();
    for (int r0 = 0; r0 < n_rows; r0 += 1024) {
        const int r = r0 + tid;
        int beg = 0, deg = 0;
        if (r < n_rows) { beg = ptr[r]; deg = ptr[r + 1] - beg; }
        const int is_long = deg > SPLIT_ROW ? 1 : 0;
        const int nseg = is_long ? (deg + SPLIT_SEG - 1) / SPLIT_SEG : 0;
        if (!__syncthreads_or(is_long)) continue;           // chunks without a split row (nearly all) cost one barrier
        int is = nseg, il = is_long;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int a = __shfl_up_sync(FULL_MASK, is, o), b = __shfl_up_sync(FULL_MASK, il, o);
            if (lane >= o) { is += a; il += b; }
        }
        if (lane == 31) { s_ws[warp] = is; s_wl[warp] = il; }
        __syncthreads();
        if (warp == 0) {
            int a = s_ws[lane], b = s_wl[lane], ia = a, ib = b;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int x = __shfl_up_sync(FULL_MASK, ia, o), y = __shfl_up_sync(FULL_MASK, ib, o);
                if (lane >= o) { ia += x; ib += y; }
            }
            s_ws[lane] = ia - a;
            s_wl[lane] = ib - b;
        }
        __syncthreads();
        const int s0 = s_base_s + s_ws[warp] + is - nseg, li = s_base_l + s_wl[warp] + il - is_long;
        if (is_long) {
            sp.long_row[li] = r;
            sp.long_seg0[li] = s0;
            sp.long_nseg[li] = nseg;
            const int per = (((deg + nseg - 1) / nseg) + 31) & ~31;
            for (int i = 0; i < nseg; ++i) {
                const int b = min(beg + deg, beg + i * per), e = min(beg + deg, b + per);
                sp.seg_row[s0 + i] = r;
                sp.seg_beg[s0 + i] = b;
                sp.seg_end[s0 + i] = e;
                sp.seg_long[s0 + i] = li;
            }
        }
        __syncthreads();
        if (tid == 1023) { s_base_s = s0 + nseg; s_base_l = li + is_long; }
        __syncthreads();
    }
    if (tid == 0) { counts[0] = s_base_s; counts[1] = s_base_l; }
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" size_t kgat_graph_split_ints(int64_t n_edges) {
    if (n_edges < 0) return 0;
    return static_cast<size_t>(SPLIT_HEADER + 2 * split_side_ints(n_edges));
}

extern "C" int kgat_graph_split_build(const kgat_graph_t* g, int32_t* counts_out_host, kgat_stream_t stream) {
    KGAT_REQUIRE(g && counts_out_host && g->split && g->csr_ptr && g->csc_ptr);
    KGAT_REQUIRE(g->n_nodes > 0 && g->n_nodes < (int64_t(1) << 31) && g->n_edges >= 0);
    cudaStream_t st = as_stream(stream);
    KGAT_CUDA(cudaMemsetAsync(g->split, 0, sizeof(int32_t) * SPLIT_HEADER, st));
    split_table_kernel<<<1, 1024, 0, st>>>(g->csr_ptr, static_cast<int>(g->n_nodes), split_side(g, 0), g->split);
    KGAT_LAUNCH_CHECK();
    split_table_kernel<<<1, 1024, 0, st>>>(g->csc_ptr, static_cast<int>(g->n_nodes), split_side(g, 1), g->split + 2);
    KGAT_LAUNCH_CHECK();
    KGAT_CUDA(cudaMemcpyAsync(counts_out_host, g->split, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    KGAT_CUDA(cudaStreamSynchronize(st));
    return KGAT_OK;
}

extern "C" size_t kgat_graph_build_workspace_bytes(int64_t n_nodes, int64_t n_edges, int64_t n_rel) {
    (void)n_nodes; (void)n_rel;
    if (n_edges < 0) return 0;
    return plan(n_edges).total;
}

extern "C" int kgat_graph_build(const kgat_graph_t* g, const int32_t* head, const int32_t* rel,
                                const int32_t* tail, void* workspace, size_t workspace_bytes,
                                kgat_stream_t stream) {
    KGAT_REQUIRE(g != nullptr);
    KGAT_REQUIRE(g->n_nodes > 0 && g->n_rel > 0 && g->n_edges >= 0);
    KGAT_REQUIRE(g->n_edges < (int64_t(1) << 31) && g->n_nodes < (int64_t(1) << 31));
    KGAT_REQUIRE(g->csr_ptr && g->csc_ptr && g->rel_ptr);
    cudaStream_t st = as_stream(stream);
    const int64_t N = g->n_nodes, E = g->n_edges, R = g->n_rel;
    if (E == 0) {
        KGAT_CUDA(cudaMemsetAsync(g->csr_ptr, 0, sizeof(int32_t) * (N + 1), st));
        KGAT_CUDA(cudaMemsetAsync(g->csc_ptr, 0, sizeof(int32_t) * (N + 1), st));
        KGAT_CUDA(cudaMemsetAsync(g->rel_ptr, 0, sizeof(int32_t) * (R + 1), st));
        return KGAT_OK;
    }
    KGAT_REQUIRE(head && rel && tail && workspace);
    KGAT_REQUIRE(g->csr_src && g->csr_dst && g->csr_rel && g->csr_eid && g->csc_dst && g->csc_pos &&
                 g->csc_eid && g->bk_pos && g->bk_head && g->bk_tail);
    Workspace w = plan(E);
    if (workspace_bytes < w.total) return KGAT_ERR_WORKSPACE;
    char* ws = static_cast<char*>(workspace);

    // CSR: stable sort of edge ids by key head * R + rel
    int rc;
    if (static_cast<uint64_t>(N) * static_cast<uint64_t>(R) < (uint64_t(1) << 32))
        rc = sort_edges_by_head_rel<uint32_t>(g, head, rel, ws, w, st);
    else
        rc = sort_edges_by_head_rel<uint64_t>(g, head, rel, ws, w, st);
    if (rc != KGAT_OK) return rc;
    gather_csr<<<blocks_for(E), TPB, 0, st>>>(head, rel, tail, g->csr_eid, E, g->csr_src, g->csr_dst, g->csr_rel);
    KGAT_LAUNCH_CHECK();
    segment_offsets<int32_t><<<blocks_for(N + 1), TPB, 0, st>>>(g->csr_dst, E, N, g->csr_ptr);
    KGAT_LAUNCH_CHECK();

    uint32_t* ka = reinterpret_cast<uint32_t*>(ws + w.keys_a);
    uint32_t* kb = reinterpret_cast<uint32_t*>(ws + w.keys_b);
    int32_t* va = reinterpret_cast<int32_t*>(ws + w.vals_a);
    int32_t* vb = reinterpret_cast<int32_t*>(ws + w.vals_b);
    uint32_t* kc = reinterpret_cast<uint32_t*>(ws + w.keys_c);          // sorted keys
    int32_t* hist = reinterpret_cast<int32_t*>(ws + w.hist);

    // CSC: stable sort of CSR positions by tail
    copy_keys_iota<<<blocks_for(E), TPB, 0, st>>>(g->csr_src, E, ka, va);
    KGAT_LAUNCH_CHECK();
    rc = radix_sort_pairs<uint32_t>(ka, kb, va, vb, kc, g->csc_pos, E, bits_for(static_cast<uint64_t>(N)), hist, st);
    if (rc != KGAT_OK) return rc;
    segment_offsets<uint32_t><<<blocks_for(N + 1), TPB, 0, st>>>(kc, E, N, g->csc_ptr);
    KGAT_LAUNCH_CHECK();
    gather2<<<blocks_for(E), TPB, 0, st>>>(g->csc_pos, E, g->csr_dst, g->csr_eid, g->csc_dst, g->csc_eid);
    KGAT_LAUNCH_CHECK();

    // relation buckets: stable sort of CSR positions by relation
    copy_keys_iota<<<blocks_for(E), TPB, 0, st>>>(g->csr_rel, E, ka, va);
    KGAT_LAUNCH_CHECK();
    rc = radix_sort_pairs<uint32_t>(ka, kb, va, vb, kc, g->bk_pos, E, bits_for(static_cast<uint64_t>(R)), hist, st);
    if (rc != KGAT_OK) return rc;
    segment_offsets<uint32_t><<<blocks_for(R + 1), TPB, 0, st>>>(kc, E, R, g->rel_ptr);
    KGAT_LAUNCH_CHECK();
    gather2<<<blocks_for(E), TPB, 0, st>>>(g->bk_pos, E, g->csr_dst, g->csr_src, g->bk_head, g->bk_tail);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_version(void) { return 1; }

extern "C" int kgat_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" const char* kgat_error_string(int status) {
    switch (status) {
        case KGAT_OK: return "ok";
        case KGAT_ERR_INVALID_ARG: return "invalid argument (null pointer, bad size or misaligned buffer)";
        case KGAT_ERR_UNSUPPORTED_DIM: return "unsupported feature width (must be a multiple of 4, at most 256)";
        case KGAT_ERR_WORKSPACE: return "workspace too small";
        case KGAT_ERR_CUDA: return "CUDA runtime error (see kgat_last_cuda_error)";
        case KGAT_ERR_NO_DEVICE: return "no sm_100 CUDA device is current";
        default: return "unknown kgat status";
    }
}

extern "C" int kgat_device_check(void) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return KGAT_ERR_NO_DEVICE; }
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
        cudaGetLastError();
        return KGAT_ERR_NO_DEVICE;
    }
    return major == 10 ? KGAT_OK : KGAT_ERR_NO_DEVICE;
}
