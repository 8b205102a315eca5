// (2) TransR attention logits and (3a/3b) segment softmax + edge-weight re-layout.
//
// Logits: pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r)   (reference models.py:135-144), computed
// relation-grouped: one CTA owns a tile of TILE edges of ONE relation bucket, stages W_r and the
// gathered e_t / e_h rows in shared memory (128-bit coalesced row loads), runs both projections as
// a register-tiled fp32 GEMM and finishes with the fused tanh / dot epilogue, so neither projection
// ever reaches HBM.  fp32 FFMA throughout: the parity bar is 1e-5 relative against the fp32/fp64
// reference, which a single TF32/BF16 tensor-core pass cannot meet.
//
// Softmax: exp(x - max) / sum over the in-edges of a destination row like DGL's edge_softmax (reference
// models.py:153), with the parallelism matched to the row length (see segment_softmax_kernel).
#include <cstdint>

#include "peer.cuh"
#include "split.cuh"

namespace kgat {
namespace {

// ------------------------------------------------------------------------------------------------
// relation-tile lookup: block b -> (relation, [e0, e1)) with tiles of TILE edges per relation bucket
// ------------------------------------------------------------------------------------------------
template <int TILE>
__device__ __forceinline__ void find_tile(const int32_t* __restrict__ rel_ptr, int n_rel, int tile,
                                          int* s_info /* [3] shared: rel, e0, e1 */) {
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 32) {
        int base = 0;
        bool found = false;
        for (int r0 = 0; r0 < n_rel; r0 += 32) {
            const int r = r0 + lane;
            const int lo = r < n_rel ? rel_ptr[r] : 0;
            const int hi = r < n_rel ? rel_ptr[r + 1] : 0;
            const int nt = (hi - lo + TILE - 1) / TILE;
            int incl = nt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            const int excl = base + incl - nt;
            const bool mine = tile >= excl && tile < excl + nt;
            const unsigned m = __ballot_sync(FULL_MASK, mine);
            if (m != 0) {
                if (mine) {
                    s_info[0] = r;
                    s_info[1] = lo + (tile - excl) * TILE;
                    s_info[2] = min(hi, lo + (tile - excl + 1) * TILE);
                }
                found = true;
                break;
            }
            base += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (!found && lane == 0) s_info[0] = -1;
    }
}

// ------------------------------------------------------------------------------------------------
// logits, D = RD = 64 fast path.  256 threads, TILE = 128 edges, each thread 8 edges x 4 columns
// for both projections.
// ------------------------------------------------------------------------------------------------
constexpr int ATT_TILE = 128;
constexpr int ATT_THREADS = 256;

template <int D, int RD>
struct AttSmem {
    static constexpr int LDA = D + 4;   // +4 floats keeps 16B alignment, spreads rows over banks
    float W[D * RD];
    float er[RD];
    float T[ATT_TILE * LDA];
    float H[ATT_TILE * LDA];
    int tail[ATT_TILE];
    int head[ATT_TILE];
    int pos[ATT_TILE];
    int info[4];
};

template <int D, int RD>
__global__ void __launch_bounds__(ATT_THREADS)
att_logits_tile_kernel(const int32_t* __restrict__ rel_ptr, int n_rel,
                       const int32_t* __restrict__ bk_head, const int32_t* __restrict__ bk_tail,
                       const int32_t* __restrict__ bk_pos, const float* __restrict__ ent,
                       const float* __restrict__ rel_emb, const float* __restrict__ W_R,
                       float* __restrict__ logits_csr) {
    static_assert(RD == 64 && D % 4 == 0, "tile kernel is specialised for relation_dim 64");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using S = AttSmem<D, RD>;
    S& s = *reinterpret_cast<S*>(smem_raw);
    constexpr int LDA = S::LDA;
    const int tid = threadIdx.x;

    find_tile<ATT_TILE>(rel_ptr, n_rel, blockIdx.x, s.info);
    __syncthreads();
    const int r = s.info[0];
    if (r < 0) return;
    const int e0 = s.info[1], e1 = s.info[2];
    const int n_e = e1 - e0;

    if (tid < ATT_TILE) {
        const bool ok = tid < n_e;
        s.tail[tid] = ok ? bk_tail[e0 + tid] : -1;
        s.head[tid] = ok ? bk_head[e0 + tid] : -1;
        s.pos[tid] = ok ? bk_pos[e0 + tid] : -1;
    }
    // W_r [D][RD] and e_r
    const float4* Wg = reinterpret_cast<const float4*>(W_R + static_cast<size_t>(r) * D * RD);
    for (int i = tid; i < D * RD / 4; i += ATT_THREADS) reinterpret_cast<float4*>(s.W)[i] = __ldg(Wg + i);
    if (tid < RD) s.er[tid] = rel_emb[static_cast<size_t>(r) * RD + tid];
    __syncthreads();

    // gather e_t / e_h rows: D/4 consecutive threads fetch one row with 128-bit loads
    constexpr int VPR = D / 4;
    for (int i = tid; i < ATT_TILE * VPR; i += ATT_THREADS) {
        const int row = i / VPR, v = i % VPR;
        const int t = s.tail[row], h = s.head[row];
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (t >= 0) {
            a = ldg4(ent + static_cast<size_t>(t) * D + v * 4);
            b = ldg4(ent + static_cast<size_t>(h) * D + v * 4);
        }
        *reinterpret_cast<float4*>(&s.T[row * LDA + v * 4]) = a;
        *reinterpret_cast<float4*>(&s.H[row * LDA + v * 4]) = b;
    }
    __syncthreads();

    const int tx = tid & 15;        // column group: columns tx*4 .. tx*4+3
    const int ty = tid >> 4;        // row group: rows ty*8 .. ty*8+7
    float accT[8][4], accH[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) accT[i][j] = accH[i][j] = 0.f;

#pragma unroll 2
    for (int k = 0; k < D; k += 4) {
        float4 b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) b[q] = *reinterpret_cast<const float4*>(&s.W[(k + q) * RD + tx * 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 at = *reinterpret_cast<const float4*>(&s.T[(ty * 8 + i) * LDA + k]);
            const float4 ah = *reinterpret_cast<const float4*>(&s.H[(ty * 8 + i) * LDA + k]);
            const float atv[4] = {at.x, at.y, at.z, at.w};
            const float ahv[4] = {ah.x, ah.y, ah.z, ah.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                accT[i][0] = fmaf(atv[q], b[q].x, accT[i][0]);
                accT[i][1] = fmaf(atv[q], b[q].y, accT[i][1]);
                accT[i][2] = fmaf(atv[q], b[q].z, accT[i][2]);
                accT[i][3] = fmaf(atv[q], b[q].w, accT[i][3]);
                accH[i][0] = fmaf(ahv[q], b[q].x, accH[i][0]);
                accH[i][1] = fmaf(ahv[q], b[q].y, accH[i][1]);
                accH[i][2] = fmaf(ahv[q], b[q].z, accH[i][2]);
                accH[i][3] = fmaf(ahv[q], b[q].w, accH[i][3]);
            }
        }
    }

    // fused epilogue: logit = sum_c t_r[c] * tanh(h_r[c] + e_r[c]); reduce over the 16 column groups
    const float4 er = *reinterpret_cast<const float4*>(&s.er[tx * 4]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float p = accT[i][0] * tanhf(accH[i][0] + er.x);
        p = fmaf(accT[i][1], tanhf(accH[i][1] + er.y), p);
        p = fmaf(accT[i][2], tanhf(accH[i][2] + er.z), p);
        p = fmaf(accT[i][3], tanhf(accH[i][3] + er.w), p);
        p = group_sum<16>(p);
        const int row = ty * 8 + i;
        if (tx == 0 && row < n_e) logits_csr[s.pos[row]] = p;
    }
}

// generic widths: one warp per edge, lane c strides over the relation_dim columns (slow, correct)
__global__ void att_logits_generic_kernel(const int32_t* __restrict__ csr_src, const int32_t* __restrict__ csr_dst,
                                          const int32_t* __restrict__ csr_rel, int64_t n_edges,
                                          const float* __restrict__ ent, const float* __restrict__ rel_emb,
                                          const float* __restrict__ W_R, int d, int rd,
                                          float* __restrict__ logits_csr) {
    const int64_t warp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_edges) return;
    const int t = csr_src[warp], h = csr_dst[warp], r = csr_rel[warp];
    const float* W = W_R + static_cast<size_t>(r) * d * rd;
    const float* et = ent + static_cast<size_t>(t) * d;
    const float* eh = ent + static_cast<size_t>(h) * d;
    float p = 0.f;
    for (int c = lane; c < rd; c += 32) {
        float tr = 0.f, hr = 0.f;
        for (int k = 0; k < d; ++k) {
            const float w = __ldg(W + static_cast<size_t>(k) * rd + c);
            tr = fmaf(__ldg(et + k), w, tr);
            hr = fmaf(__ldg(eh + k), w, hr);
        }
        p = fmaf(tr, tanhf(hr + __ldg(rel_emb + static_cast<size_t>(r) * rd + c)), p);
    }
    p = warp_sum(p);
    if (lane == 0) logits_csr[warp] = p;
}

// ------------------------------------------------------------------------------------------------
// segment softmax: warp per destination row
// ------------------------------------------------------------------------------------------------
// Segment softmax (reference models.py:153, DGL edge_softmax: exp(x - max) / sum over the in-edges of a row).
// Rows come at three sizes and each size gets the parallelism that fits it:
//   deg <= 32           eight lanes per row, four rows per warp (44 % of the rows of a CKG are users with ~8
//                       interactions: a full warp per row left three quarters of the lanes idle); the <= 4
//                       logits of a lane stay in registers, max and sum by 8-lane shuffles - one pass over memory;
//   32 < deg            the whole warp, ONLINE max / sum (one pass: a lane keeps a running maximum m and the sum
//                       of exp(x - m), rescaled when m grows; lanes are merged by shuffles), then the write pass;
//   deg > SPLIT_ROW     (split.cuh) one CTA per segment: segment (max, sum) -> per-row merge -> write, three small
//                       kernels that run only when the graph has such rows.
// PUSH (multi-GPU): every weight is also stored to the peers' copies of the all-edges vector (peer.cuh); the
// single-GPU instantiation carries no trace of it.
constexpr int SM_ROWS_PER_WARP = 4;
constexpr int SM_ROWS_PER_CTA = 8 * SM_ROWS_PER_WARP;

template <bool PUSH>
__global__ void __launch_bounds__(256)
segment_softmax_kernel(const int32_t* __restrict__ csr_ptr, const int32_t* __restrict__ csr_eid,
                       int64_t row_begin, int64_t row_end, const float* __restrict__ logits,
                       float* __restrict__ w_csr, float* __restrict__ w_eid, int split_done, PeerOut po) {
    const int lane = threadIdx.x & 31, sub = lane >> 3, sl = lane & 7;
    const int64_t warp_row = row_begin + (blockIdx.x * 8 + (threadIdx.x >> 5)) * static_cast<int64_t>(SM_ROWS_PER_WARP);
    if (warp_row >= row_end) return;
    const int64_t row = warp_row + sub;
    int beg = 0, end = 0;
    if (row < row_end) { beg = __ldg(csr_ptr + row); end = __ldg(csr_ptr + row + 1); }
    const int deg = end - beg;
    if (deg <= 32) {
        // the branch is uniform over the 8 lanes of a row; the shuffles name exactly those lanes
        const unsigned gmask = 0xffu << (sub * 8);
        float x[4];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = beg + sl + 8 * j;
            x[j] = c < end ? __ldg(logits + c) : -INFINITY;
            m = fmaxf(m, x[j]);
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(gmask, m, o));
        float den = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            x[j] = beg + sl + 8 * j < end ? expf(x[j] - m) : 0.f;
            den += x[j];
        }
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) den += __shfl_xor_sync(gmask, den, o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = beg + sl + 8 * j;
            if (c < end) {
                const float w = x[j] / den;
                w_csr[c] = w;
                if (PUSH) peer_store(po, c, w);
                if (w_eid) w_eid[__ldg(csr_eid + c)] = w;
            }
        }
    }
    // rows above 32 edges: one after the other, by the whole warp
    unsigned todo = __ballot_sync(FULL_MASK, deg > 32 && sl == 0);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int b = __shfl_sync(FULL_MASK, beg, src), e = __shfl_sync(FULL_MASK, end, src);
        if (split_done && e - b > SPLIT_ROW) continue;        // the segment kernels own this row
        float m = -INFINITY, sum = 0.f;
        for (int c = b + lane; c < e; c += 32) {
            const float x = __ldg(logits + c);
            if (x > m) {                    // exp(-inf) = 0 covers the first element
                sum = sum * expf(m - x) + 1.f;
                m = x;
            } else {
                sum += expf(x - m);
            }
        }
        const float M = warp_max(m);        // every lane holds at least one element (deg > 32): m is finite
        const float den = warp_sum(sum * expf(m - M));
        for (int c = b + lane; c < e; c += 32) {
            const float w = expf(__ldg(logits + c) - M) / den;
            w_csr[c] = w;
            if (PUSH) peer_store(po, c, w);
            if (w_eid) w_eid[__ldg(csr_eid + c)] = w;
        }
    }
}

// one CTA per segment of a split row: (max, sum of exp(x - max)) of the segment
__global__ void __launch_bounds__(256)
softmax_seg_stats_kernel(SplitSide sp, const float* __restrict__ logits, int64_t row_begin, int64_t row_end,
                         float2* __restrict__ stats) {
    __shared__ float s_red[8];
    const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = __ldg(sp.seg_row + s);
    if (row < row_begin || row >= row_end) return;
    const int beg = __ldg(sp.seg_beg + s), end = __ldg(sp.seg_end + s);
    float m = -INFINITY;
    for (int c = beg + threadIdx.x; c < end; c += 256) m = fmaxf(m, __ldg(logits + c));
    m = warp_max(m);
    if (lane == 0) s_red[warp] = m;
    __syncthreads();
    float M = s_red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) M = fmaxf(M, s_red[w]);
    __syncthreads();
    float sum = 0.f;
    for (int c = beg + threadIdx.x; c < end; c += 256) sum += expf(__ldg(logits + c) - M);
    sum = warp_sum(sum);
    if (lane == 0) s_red[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = s_red[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += s_red[w];
        stats[s] = make_float2(M, t);
    }
}

// one CTA per segment: merge the row's segment statistics (every CTA of the row repeats the same fixed-order
// merge, a few hundred values at most), then write this segment's weights
template <bool PUSH>
__global__ void __launch_bounds__(256)
softmax_seg_write_kernel(SplitSide sp, const float2* __restrict__ stats, const float* __restrict__ logits,
                         const int32_t* __restrict__ csr_eid, int64_t row_begin, int64_t row_end,
                         float* __restrict__ w_csr, float* __restrict__ w_eid, PeerOut po) {
    __shared__ float s_M, s_den;
    const int s = blockIdx.x, lane = threadIdx.x & 31;
    const int row = __ldg(sp.seg_row + s);
    if (row < row_begin || row >= row_end) return;
    if (threadIdx.x < 32) {
        const int li = __ldg(sp.seg_long + s);
        const int s0 = __ldg(sp.long_seg0 + li), n = __ldg(sp.long_nseg + li);
        float m = -INFINITY;
        for (int i = lane; i < n; i += 32) m = fmaxf(m, stats[s0 + i].x);
        const float M = warp_max(m);
        float t = 0.f;
        for (int i = lane; i < n; i += 32) {
            const float2 q = stats[s0 + i];
            t += q.y * expf(q.x - M);
        }
        t = warp_sum(t);
        if (lane == 0) { s_M = M; s_den = t; }
    }
    __syncthreads();
    const float M = s_M, den = s_den;
    const int beg = __ldg(sp.seg_beg + s), end = __ldg(sp.seg_end + s);
    for (int c = beg + threadIdx.x; c < end; c += 256) {
        const float w = expf(__ldg(logits + c) - M) / den;
        w_csr[c] = w;
        if (PUSH) peer_store(po, c, w);
        if (w_eid) w_eid[__ldg(csr_eid + c)] = w;
    }
}

__global__ void gather_weights_kernel(const float* __restrict__ w_eid, const int32_t* __restrict__ eid,
                                      int64_t n, float* __restrict__ out) {
    int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) out[i] = __ldg(w_eid + eid[i]);
}

}  // namespace
}  // namespace kgat

using namespace kgat;

#include <cstdlib>
#include <cstring>

int kgat_launch_att_logits_umma(const kgat_graph_t* g, const float* entity, const float* rel_emb, const float* W_R,
                                float* logits_csr, int* only_if, cudaStream_t st);
int kgat_launch_att_logits_umma16(const kgat_graph_t* g, const float* entity, const float* rel_emb,
                                  const float* W_R, float* logits_csr, int* fallback, const int32_t* rel_skip,
                                  cudaStream_t st);
int kgat_launch_att_range(const float* W_R, int64_t n, int* word, cudaStream_t st);
int kgat_launch_att_logits_runs(const kgat_graph_t* g, const kgat_runs_t* runs, const float* entity,
                                const float* rel_emb, const float* W_R, float* logits_csr, float* v_rows,
                                int64_t v_base, int64_t v_count, int64_t p_begin, int64_t p_end,
                                const int32_t* rel_mode, int* fallback, cudaStream_t st);
int* kgat_att_fallback_word();

// KGAT_ATT_IMPL selects the kernel for d = rd = 64 (A/B runs): "f16" (default) tcgen05 split-fp16,
// "tf32" tcgen05 3xTF32, "ffma" fp32 FFMA tiles.
static int att_impl() {
    const char* e = std::getenv("KGAT_ATT_IMPL");
    if (e && std::strcmp(e, "ffma") == 0) return 0;
    if (e && std::strcmp(e, "tf32") == 0) return 1;
    return 2;
}

extern "C" int kgat_attention_logits(const kgat_graph_t* g, const float* entity, const float* rel_emb,
                                     const float* W_R, int d, int rd, float* logits_csr,
                                     kgat_stream_t stream) {
    KGAT_REQUIRE(g && entity && rel_emb && W_R && logits_csr);
    KGAT_REQUIRE(aligned16(entity) && aligned16(rel_emb) && aligned16(W_R));
    if (d <= 0 || rd <= 0 || d % 4 != 0 || rd % 4 != 0) return KGAT_ERR_UNSUPPORTED_DIM;
    if (g->n_edges == 0) return KGAT_OK;
    cudaStream_t st = as_stream(stream);
    const int64_t E = g->n_edges;
    const int impl = att_impl();
    if (d == 64 && rd == 64 && g->n_rel <= 1024 && impl == 2) {
        // The split-fp16 kernel is exact to ~3e-7 only for |W_r| < 32 and |entity| < 65504; the reference
        // computes in fp32 (models.py:140-143).  Three launches, no host round trip (capture-safe): the range
        // kernel decides from W_R, the fp16 kernel runs unless told not to and reports non-finite logits, the
        // 3xTF32 kernel (fp32 range) recomputes everything iff the word is set - an empty launch otherwise.
        int* word = kgat_att_fallback_word();
        if (!word) return KGAT_ERR_CUDA;
        int rc = kgat_launch_att_range(W_R, g->n_rel * static_cast<int64_t>(d) * rd, word, st);
        if (rc != KGAT_OK) return rc;
        rc = kgat_launch_att_logits_umma16(g, entity, rel_emb, W_R, logits_csr, word, nullptr, st);
        if (rc != KGAT_OK) return rc;
        return kgat_launch_att_logits_umma(g, entity, rel_emb, W_R, logits_csr, word, st);
    } else if (d == 64 && rd == 64 && g->n_rel <= 1024 && impl == 1) {
        return kgat_launch_att_logits_umma(g, entity, rel_emb, W_R, logits_csr, nullptr, st);
    } else if (d == 64 && rd == 64) {
        using S = AttSmem<64, 64>;
        static PerDeviceOnce configured;
        if (configured.need()) {
            KGAT_CUDA(cudaFuncSetAttribute(att_logits_tile_kernel<64, 64>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(S)));
            configured.set();
        }
        // tiles = sum_r ceil(E_r / TILE) <= E / TILE + R; surplus blocks exit immediately
        const int64_t grid = E / ATT_TILE + g->n_rel;
        att_logits_tile_kernel<64, 64><<<static_cast<unsigned>(grid), ATT_THREADS, sizeof(S), st>>>(
            g->rel_ptr, static_cast<int>(g->n_rel), g->bk_head, g->bk_tail, g->bk_pos, entity, rel_emb, W_R,
            logits_csr);
    } else {
        const int64_t blocks = (E * 32 + 255) / 256;
        att_logits_generic_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            g->csr_src, g->csr_dst, g->csr_rel, E, entity, rel_emb, W_R, d, rd, logits_csr);
    }
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_attention_logits_runs(const kgat_graph_t* g, const kgat_runs_t* runs, const float* entity,
                                          const float* rel_emb, const float* W_R, int d, int rd, float* logits_csr,
                                          float* v_rows, kgat_stream_t stream) {
    KGAT_REQUIRE(g && runs && entity && rel_emb && W_R && logits_csr);
    KGAT_REQUIRE(runs->bk_run && runs->run_ptr && runs->run_head && runs->rel_run_ptr && runs->n_runs);
    KGAT_REQUIRE(aligned16(entity) && aligned16(rel_emb) && aligned16(W_R));
    if (d != 64 || rd != 64 || g->n_rel > 1024) return KGAT_ERR_UNSUPPORTED_DIM;
    if (g->n_edges == 0) return KGAT_OK;
    cudaStream_t st = as_stream(stream);
    // same three-launch range protocol as kgat_attention_logits: range word, fp16 kernel, fp32-range kernel iff needed
    int* word = kgat_att_fallback_word();
    if (!word) return KGAT_ERR_CUDA;
    int rc = kgat_launch_att_range(W_R, g->n_rel * static_cast<int64_t>(d) * rd, word, st);
    if (rc != KGAT_OK) return rc;
    KGAT_REQUIRE(!v_rows || aligned16(v_rows));
    rc = kgat_launch_att_logits_runs(g, runs, entity, rel_emb, W_R, logits_csr, v_rows, 0, INT64_MAX, 0, g->n_edges,
                                     nullptr, word, st);
    if (rc != KGAT_OK) return rc;
    return kgat_launch_att_logits_umma(g, entity, rel_emb, W_R, logits_csr, word, st);
}

extern "C" int kgat_attention_logits_hybrid(const kgat_graph_t* g, const kgat_runs_t* runs, const float* entity,
                                            const float* rel_emb, const float* W_R, int d, int rd, float* logits_csr,
                                            float* v_rows, int64_t v_base, int64_t v_count, int64_t p_begin,
                                            int64_t p_end, kgat_stream_t stream) {
    KGAT_REQUIRE(g && runs && entity && rel_emb && W_R && logits_csr);
    KGAT_REQUIRE(runs->bk_run && runs->run_ptr && runs->run_head && runs->rel_run_ptr && runs->n_runs && runs->rel_mode);
    KGAT_REQUIRE(aligned16(entity) && aligned16(rel_emb) && aligned16(W_R) && (!v_rows || aligned16(v_rows)));
    KGAT_REQUIRE(v_base >= 0 && v_count >= 0 && 0 <= p_begin && p_begin <= p_end && p_end <= g->n_edges);
    if (d != 64 || rd != 64 || g->n_rel > 1024) return KGAT_ERR_UNSUPPORTED_DIM;
    if (g->n_edges == 0) return KGAT_OK;
    cudaStream_t st = as_stream(stream);
    int* word = kgat_att_fallback_word();
    if (!word) return KGAT_ERR_CUDA;
    int rc = kgat_launch_att_range(W_R, g->n_rel * static_cast<int64_t>(d) * rd, word, st);
    if (rc != KGAT_OK) return rc;
    // relations marked 0: two projections per edge; relations marked 1: two per (head, relation) run + a dot per edge
    rc = kgat_launch_att_logits_umma16(g, entity, rel_emb, W_R, logits_csr, word, runs->rel_mode, st);
    if (rc != KGAT_OK) return rc;
    if (p_end > p_begin) {
        rc = kgat_launch_att_logits_runs(g, runs, entity, rel_emb, W_R, logits_csr, v_rows, v_base, v_count, p_begin,
                                         p_end, runs->rel_mode, word, st);
        if (rc != KGAT_OK) return rc;
    }
    return kgat_launch_att_logits_umma(g, entity, rel_emb, W_R, logits_csr, word, st);
}

extern "C" int kgat_segment_softmax_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                                         const float* logits_csr, float* w_csr, float* w_eid,
                                         const kgat_peer_group_t* pg, kgat_stream_t stream) {
    KGAT_REQUIRE(g && logits_csr && w_csr);
    KGAT_REQUIRE(row_begin >= 0 && row_end <= g->n_nodes && row_begin <= row_end);
    PeerOut po = no_peers();
    if (pg && g->n_edges > 0) KGAT_REQUIRE(peer_aliases(pg, w_csr, sizeof(float) * g->n_edges, &po));
    if (row_begin == row_end || g->n_edges == 0) return KGAT_OK;
    const int64_t rows = row_end - row_begin;
    const int64_t blocks = (rows + SM_ROWS_PER_CTA - 1) / SM_ROWS_PER_CTA;
    cudaStream_t st = as_stream(stream);
    const bool split = split_active(g, 0);
    if (pg)
        segment_softmax_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            g->csr_ptr, g->csr_eid, row_begin, row_end, logits_csr, w_csr, w_eid, split ? 1 : 0, po);
    else
        segment_softmax_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, st>>>(
            g->csr_ptr, g->csr_eid, row_begin, row_end, logits_csr, w_csr, w_eid, split ? 1 : 0, po);
    KGAT_LAUNCH_CHECK();
    if (split) {
        const SplitSide sp = split_side(g, 0);
        float2* stats = reinterpret_cast<float2*>(g->split_scratch);
        softmax_seg_stats_kernel<<<sp.n_seg, 256, 0, st>>>(sp, logits_csr, row_begin, row_end, stats);
        KGAT_LAUNCH_CHECK();
        if (pg)
            softmax_seg_write_kernel<true><<<sp.n_seg, 256, 0, st>>>(sp, stats, logits_csr, g->csr_eid, row_begin,
                                                                   row_end, w_csr, w_eid, po);
        else
            softmax_seg_write_kernel<false><<<sp.n_seg, 256, 0, st>>>(sp, stats, logits_csr, g->csr_eid, row_begin,
                                                                    row_end, w_csr, w_eid, po);
        KGAT_LAUNCH_CHECK();
    }
    return KGAT_OK;
}

extern "C" int kgat_segment_softmax(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                                    const float* logits_csr, float* w_csr, float* w_eid,
                                    kgat_stream_t stream) {
    return kgat_segment_softmax_push(g, row_begin, row_end, logits_csr, w_csr, w_eid, nullptr, stream);
}

extern "C" int kgat_edge_weights_from_eid(const kgat_graph_t* g, const float* w_eid, float* w_csr,
                                          float* w_csc, kgat_stream_t stream) {
    KGAT_REQUIRE(g);
    if (g->n_edges == 0) return KGAT_OK;
    KGAT_REQUIRE(w_eid);
    const int64_t E = g->n_edges;
    const unsigned blocks = static_cast<unsigned>((E + 255) / 256);
    if (w_csr) {
        gather_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_eid, g->csr_eid, E, w_csr);
        KGAT_LAUNCH_CHECK();
    }
    if (w_csc) {
        gather_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w_eid, g->csc_eid, E, w_csc);
        KGAT_LAUNCH_CHECK();
    }
    return KGAT_OK;
}
