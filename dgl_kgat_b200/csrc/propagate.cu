// (3c)+(4) KGATConv forward / backward and the entity-table gradient.
//
// Forward (reference models.py:60-70 + :163-167), one CTA per tile of TM destination rows:
//   phase 1  warp-per-row CSR aggregation  h_nb[v] = sum_c w[c] * h_in[src[c]]  with 128-bit row
//            loads (DIN/4 lanes per edge, 32/(DIN/4) edges per warp instruction, 4x unrolled),
//            fp32 accumulation in registers, fixed summation order (no atomics);
//            z = h_in[v] (.) h_nb[v] goes to shared memory, h_nb to HBM for the backward;
//   phase 2  tile GEMM  pre = z W2^T  from shared memory, LeakyReLU, dropout mask;
//   phase 3  row L2 norm by warp shuffle, un-normalised row -> h_out, normalised row -> its slice of
//            the (N, sum d) concat buffer.
// Backward, persistent CTAs over row tiles:
//   phase 1  gradient of the layer output assembled per row: normalize-backward of the concat slice
//            + next layer's direct term + transposed aggregation over the CSC (again gather-only,
//            deterministic), then dropout / LeakyReLU backward -> dpre tile in shared memory;
//   phase 2  dW2 += dpre^T z in registers (4xBJ blocks per thread, one partial per CTA, reduced by a
//            second kernel in fixed order), dz = dpre W2, dx_direct = dz (.) h_nb, dh_nb = dz (.) h_in.
//
// Multi-GPU (peer.cuh): the PUSH instantiations store every owned row of the exchanged output (h_out,
// dh_nb, d_entity) into the peers' arenas right next to the local store, so the all-gather of the
// destination-partitioned step rides on the kernel that produces the rows.
#include <cstdlib>
#include <cstring>

#include "peer.cuh"
#include "split.cuh"

namespace kgat {
namespace {

constexpr int TM = 32;          // destination rows per tile
constexpr int THREADS = 256;    // 8 warps
constexpr int WARPS = THREADS / 32;

// sum_c w[c] * feat[idx[c]] over c in [beg,end) for one row; W-wide rows, every lane group ends
// with the full sum (lane l holds elements 4*(l % (W/4)) .. +3)
template <int W>
__device__ __forceinline__ float4 gather_row(const int32_t* __restrict__ idx, const float* __restrict__ wgt,
                                             int beg, int end, const float* __restrict__ feat, int lane) {
    constexpr int LPE = W / 4;          // lanes per edge
    constexpr int EPW = 32 / LPE;       // edges per warp instruction
    const int sub = lane / LPE, li = lane % LPE;
    const float* base = feat + li * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = beg + sub;
    for (; e + 3 * EPW < end; e += 4 * EPW) {
        const int s0 = __ldg(idx + e), s1 = __ldg(idx + e + EPW), s2 = __ldg(idx + e + 2 * EPW),
                  s3 = __ldg(idx + e + 3 * EPW);
        const float w0 = __ldg(wgt + e), w1 = __ldg(wgt + e + EPW), w2 = __ldg(wgt + e + 2 * EPW),
                    w3 = __ldg(wgt + e + 3 * EPW);
        const float4 v0 = ldg4(base + static_cast<size_t>(s0) * W);
        const float4 v1 = ldg4(base + static_cast<size_t>(s1) * W);
        const float4 v2 = ldg4(base + static_cast<size_t>(s2) * W);
        const float4 v3 = ldg4(base + static_cast<size_t>(s3) * W);
        fma4(acc, w0, v0);
        fma4(acc, w1, v1);
        fma4(acc, w2, v2);
        fma4(acc, w3, v3);
    }
    // tail: at most three edges per lane group are left.  Their loads are issued together (a short row -
    // the ~8 interactions of a user, 44 % of the rows of a CKG - would otherwise be a chain of dependent
    // index -> feature-row round trips); the additions keep the order of the edges.
    if (e < end) {
        const bool p1 = e + EPW < end, p2 = e + 2 * EPW < end;
        const int s0 = __ldg(idx + e), s1 = p1 ? __ldg(idx + e + EPW) : 0, s2 = p2 ? __ldg(idx + e + 2 * EPW) : 0;
        const float w0 = __ldg(wgt + e), w1 = p1 ? __ldg(wgt + e + EPW) : 0.f, w2 = p2 ? __ldg(wgt + e + 2 * EPW) : 0.f;
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 v0 = ldg4(base + static_cast<size_t>(s0) * W);
        const float4 v1 = p1 ? ldg4(base + static_cast<size_t>(s1) * W) : zero;
        const float4 v2 = p2 ? ldg4(base + static_cast<size_t>(s2) * W) : zero;
        fma4(acc, w0, v0);
        if (p1) fma4(acc, w1, v1);
        if (p2) fma4(acc, w2, v2);
    }
    return across_groups_sum<LPE>(acc);
}

// Rows with more than LONG_ROW edges (hot entities of a power-law KG) are gathered by ALL warps of the
// CTA: warp w takes the w-th contiguous slice (a multiple of 32 edges, so the summation order depends
// on the row only), partials meet in shared memory and are added in warp order - deterministic.
constexpr int LONG_ROW = 512;
constexpr int COOP_FLOATS = WARPS * 32 * 4 + TM;      // float4 partial per lane per warp + row flags

// CTA-uniform call; every lane of every warp returns the total (element group lane % (W/4))
template <int W>
__device__ __forceinline__ float4 gather_row_cta(const int32_t* __restrict__ idx, const float* __restrict__ wgt,
                                                 int beg, int end, const float* __restrict__ feat, float4* sPart,
                                                 int warp, int lane) {
    const int per = (((end - beg) + WARPS - 1) / WARPS + 31) & ~31;
    const int b = min(end, beg + warp * per), e = min(end, b + per);
    sPart[warp * 32 + lane] = gather_row<W>(idx, wgt, b, e, feat, lane);
    __syncthreads();
    float4 t = sPart[lane];
#pragma unroll
    for (int w = 1; w < WARPS; ++w) {
        const float4 q = sPart[w * 32 + lane];
        t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
    }
    __syncthreads();
    return t;
}

// Rows with more than SPLIT_ROW edges (split.cuh: a few hundred edges already - hot rows come in clusters of
// neighbouring ids and would keep one CTA busy long after the grid has drained - up to the 10^6 of the hottest
// entity of a power-law graph) are gathered by a PRE-PASS in front of the kernel that owns the row: one CTA per segment of <= SPLIT_SEG
// edges writes a partial sum, one warp per split row adds the partials in segment order into the row of `out`
// where the main kernel picks the total up instead of walking the row (fixed order: deterministic).
template <int W>
__global__ void __launch_bounds__(THREADS)
split_gather_kernel(SplitSide sp, const int32_t* __restrict__ idx, const float* __restrict__ wgt,
                    const float* __restrict__ feat, int row_begin, int row_end, float* __restrict__ partial) {
    __shared__ float4 sPart[WARPS * 32];
    const int s = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = __ldg(sp.seg_row + s);
    if (row < row_begin || row >= row_end) return;            // CTA-uniform
    const float4 t = gather_row_cta<W>(idx, wgt, __ldg(sp.seg_beg + s), __ldg(sp.seg_end + s), feat, sPart, warp, lane);
    if (warp == 0 && lane < W / 4) *reinterpret_cast<float4*>(partial + static_cast<size_t>(s) * W + lane * 4) = t;
}

// One CTA per split row: the hottest row of a Zipf graph has > 1000 segments, and one warp adding them one after
// the other was a 200 us chain of dependent loads.  Warp w takes the w-th contiguous slice of the row's partials
// (32 / (W / 4) partial rows per warp instruction, 4x unrolled), the lane groups meet by shuffle, the warps in
// shared memory in warp order - the order of the additions depends on the segment count only.
template <int W>
__global__ void __launch_bounds__(THREADS)
split_combine_kernel(SplitSide sp, const float* __restrict__ partial, int row_begin, int row_end,
                     float* __restrict__ out) {
    constexpr int LPE = W / 4, EPW = 32 / LPE;
    __shared__ float4 sPart[WARPS * 32];
    const int li = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = __ldg(sp.long_row + li);
    if (row < row_begin || row >= row_end) return;            // CTA-uniform
    const int s0 = __ldg(sp.long_seg0 + li), n = __ldg(sp.long_nseg + li);
    const int per = (((n + WARPS - 1) / WARPS) + EPW - 1) / EPW * EPW;
    const int b = min(n, warp * per), e = min(n, b + per);
    const int sub = lane / LPE;
    const float* base = partial + static_cast<size_t>(s0) * W + (lane % LPE) * 4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int i = b + sub;
    for (; i + 3 * EPW < e; i += 4 * EPW) {
        const float4 v0 = ldg4(base + static_cast<size_t>(i) * W), v1 = ldg4(base + static_cast<size_t>(i + EPW) * W);
        const float4 v2 = ldg4(base + static_cast<size_t>(i + 2 * EPW) * W),
                     v3 = ldg4(base + static_cast<size_t>(i + 3 * EPW) * W);
        acc.x += v0.x; acc.y += v0.y; acc.z += v0.z; acc.w += v0.w;
        acc.x += v1.x; acc.y += v1.y; acc.z += v1.z; acc.w += v1.w;
        acc.x += v2.x; acc.y += v2.y; acc.z += v2.z; acc.w += v2.w;
        acc.x += v3.x; acc.y += v3.y; acc.z += v3.z; acc.w += v3.w;
    }
    for (; i < e; i += EPW) {
        const float4 v = ldg4(base + static_cast<size_t>(i) * W);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    sPart[warp * 32 + lane] = across_groups_sum<LPE>(acc);
    __syncthreads();
    if (warp == 0 && lane < LPE) {
        float4 t = sPart[lane];
#pragma unroll
        for (int w = 1; w < WARPS; ++w) {
            const float4 q = sPart[w * 32 + lane];
            t.x += q.x; t.y += q.y; t.z += q.z; t.w += q.w;
        }
        *reinterpret_cast<float4*>(out + static_cast<size_t>(row) * W + lane * 4) = t;
    }
}

// side 0: CSR (idx = csr_src), side 1: CSC (idx = csc_dst).  Returns true when the rows above SPLIT_ROW edges of
// [row_begin, row_end) now hold their gathered sums in `out` ((n_nodes, W) row-major).
template <int W>
int split_prepass(const kgat_graph_t* g, int side, const float* wgt, const float* feat, int row_begin, int row_end,
                  float* out, cudaStream_t st, bool* ran) {
    *ran = false;
    if (!split_active(g, side)) return KGAT_OK;
    static_assert(W <= KGAT_SPLIT_SCRATCH_WIDTH, "scratch rows are KGAT_SPLIT_SCRATCH_WIDTH floats");
    const SplitSide sp = split_side(g, side);
    const int32_t* idx = side == 0 ? g->csr_src : g->csc_dst;
    split_gather_kernel<W><<<sp.n_seg, THREADS, 0, st>>>(sp, idx, wgt, feat, row_begin, row_end, g->split_scratch);
    KGAT_LAUNCH_CHECK();
    split_combine_kernel<W><<<sp.n_long, THREADS, 0, st>>>(sp, g->split_scratch, row_begin, row_end, out);
    KGAT_LAUNCH_CHECK();
    *ran = true;
    return KGAT_OK;
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
// EPI selects the tile GEMM of phase 2: 0 = fp32 FFMA from shared memory (sW transposed, [DIN][DOUT+1]);
// 1 = tensor cores, mma.sync m16n8k8 3xTF32 (x = hi + lo; lo*hi + hi*lo + hi*hi, fp32 accumulate).  For the
// MMA the operands are read as fragments straight from shared memory: sZ rows and W2 rows (W2 as given,
// [DOUT][DIN] = the "col" B operand) both with a row stride of DIN + 4 floats, which makes every fragment
// load (lane = (g, t) reads row g, column t) hit 32 different banks.
template <int DIN, int DOUT, int EPI = 0>
struct FwdSmem {
    static constexpr int LDZ = EPI ? DIN + 4 : DIN;
    static constexpr int LDW = EPI ? DIN + 4 : DOUT + 1;
    static constexpr int WROWS = EPI ? DOUT : DIN;
    static constexpr int LDY = EPI ? DOUT + 8 : DOUT;
    // the long-row scratch (phase 1 only) may live in the sY region (phase 2 onwards) when that is big enough
    static constexpr bool ALIAS = EPI == 1 && TM * LDY >= COOP_FLOATS;
    static constexpr size_t bytes =
        sizeof(float) * (WROWS * LDW + TM * LDZ + TM * LDY + (ALIAS ? 0 : COOP_FLOATS)) + 16;
};

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    // hi = round-to-nearest TF32 of x (bit trick of umma.cuh), lo = x - hi (exact; the tensor core truncates it)
    hi = (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int DIN, int DOUT, bool PUSH, int EPI>
__global__ void __launch_bounds__(THREADS)
layer_fwd_kernel(const int32_t* __restrict__ csr_ptr, const int32_t* __restrict__ csr_src,
                 const float* __restrict__ w_csr, int row_begin, int row_end,
                 const float* __restrict__ h_in, const float* __restrict__ W2,
                 const float* __restrict__ drop_mask, float slope, float* __restrict__ h_nb,
                 float* __restrict__ act, float* __restrict__ h_out, float* __restrict__ concat,
                 int64_t cstride, int64_t out_off, int64_t in_off, int split_done, PeerOut po) {
    using S = FwdSmem<DIN, DOUT, EPI>;
    constexpr int LDW = S::LDW, LDZ = S::LDZ, LDY = S::LDY;
    constexpr int LPI = DIN / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sZ = reinterpret_cast<float*>(smem_raw);          // [TM][LDZ]   (16B aligned)
    float* sY = sZ + TM * LDZ;                               // [TM][LDY]
    float* sW = sY + TM * LDY;                               // EPI 0: [DIN][LDW], sW[i][j] = W2[j][i];  EPI 1: W2 as given, [DOUT][LDW]
    float4* sPart = S::ALIAS ? reinterpret_cast<float4*>(sY)
                             : reinterpret_cast<float4*>(sW + S::WROWS * LDW + ((4 - (S::WROWS * LDW) % 4) % 4));   // 16B aligned
    int* sLong = reinterpret_cast<int*>(sPart + WARPS * 32);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row0 = row_begin + blockIdx.x * TM;

    if (EPI == 0) {
        for (int idx = tid; idx < DOUT * DIN; idx += THREADS) {
            const int j = idx / DIN, i = idx % DIN;
            sW[i * LDW + j] = __ldg(W2 + idx);
        }
    } else {
        for (int idx = tid; idx < DOUT * DIN / 4; idx += THREADS) {
            const int j = idx / LPI, i4 = idx % LPI;
            *reinterpret_cast<float4*>(sW + j * LDW + i4 * 4) = __ldg(reinterpret_cast<const float4*>(W2) + idx);
        }
    }

    // phase 1: aggregation, z tile
    auto finish_row = [&](int row, int rl, const float4& hn) {
        if (lane < LPI) {
            const size_t o = static_cast<size_t>(row) * DIN + lane * 4;
            const float4 x = ldg4(h_in + o);
            *reinterpret_cast<float4*>(h_nb + o) = hn;
            if (in_off >= 0)
                *reinterpret_cast<float4*>(concat + static_cast<size_t>(row) * cstride + in_off + lane * 4) = x;
            *reinterpret_cast<float4*>(sZ + rl * LDZ + lane * 4) =
                make_float4(x.x * hn.x, x.y * hn.y, x.z * hn.z, x.w * hn.w);
        }
    };
    bool warp_long = false;
    for (int rl = warp; rl < TM; rl += WARPS) {
        const int row = row0 + rl;
        bool is_long = false;
        if (row < row_end) {
            const int beg = __ldg(csr_ptr + row), end = __ldg(csr_ptr + row + 1);
            // with the pre-pass every row above SPLIT_ROW edges is already gathered; without it rows above LONG_ROW
            // go to the whole CTA
            if (end - beg > (split_done ? SPLIT_ROW : LONG_ROW)) is_long = true;
            else finish_row(row, rl, gather_row<DIN>(csr_src, w_csr, beg, end, h_in, lane));
        } else if (lane < LPI) {
            *reinterpret_cast<float4*>(sZ + rl * LDZ + lane * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (lane == 0) sLong[rl] = is_long ? 1 : 0;
        warp_long |= is_long;
    }
    // the barrier that ends phase 1 also tells every thread whether the tile holds a long row
    if (__syncthreads_or(warp_long ? 1 : 0)) {
        for (int rl = 0; rl < TM; ++rl) {
            if (!sLong[rl]) continue;                     // CTA-uniform
            const int row = row0 + rl;
            const int beg = __ldg(csr_ptr + row), end = __ldg(csr_ptr + row + 1);
            if (split_done && end - beg > SPLIT_ROW) {    // gathered by the pre-pass (split_prepass) into h_nb
                if (warp == rl % WARPS)
                    finish_row(row, rl, *reinterpret_cast<const float4*>(h_nb + static_cast<size_t>(row) * DIN +
                                                                         (lane % LPI) * 4));
                continue;
            }
            const float4 hn = gather_row_cta<DIN>(csr_src, w_csr, beg, end, h_in, sPart, warp, lane);
            if (warp == rl % WARPS) finish_row(row, rl, hn);
        }
        __syncthreads();
    }

    // phase 2: pre = z W2^T, LeakyReLU, dropout
    auto emit = [&](int rl, int j, float pre) {           // activation of one output element -> act, sY
        const int row = row0 + rl;
        float y = 0.f;
        if (row < row_end) {
            const float a = pre > 0.f ? pre : pre * slope;
            const size_t o = static_cast<size_t>(row) * DOUT + j;
            if (act) act[o] = a;
            y = drop_mask ? a * __ldg(drop_mask + o) : a;
        }
        sY[rl * LDY + j] = y;
    };
    if (EPI == 1) {
        static_assert(TM == 32 && DIN % 8 == 0 && DOUT % 8 == 0, "m16n8k8 tiling");
        constexpr int NT = DOUT / 8, KS = DIN / 8;
        const int g = lane >> 2, t = lane & 3;
        for (int p = warp; p < 2 * NT; p += WARPS) {      // (16-row block, 8-column block) pairs of the 32 x DOUT tile
            const int mt = p / NT, nt = p % NT;
            const float* z0 = sZ + (16 * mt + g) * LDZ + t;       // A rows g, g + 8
            const float* z1 = z0 + 8 * LDZ;
            const float* wr = sW + (8 * nt + g) * LDW + t;        // B column g = row g of W2
            float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                uint32_t ah[4], al[4], bh[2], bl[2];
                split_tf32(z0[8 * ks], ah[0], al[0]);
                split_tf32(z1[8 * ks], ah[1], al[1]);
                split_tf32(z0[8 * ks + 4], ah[2], al[2]);
                split_tf32(z1[8 * ks + 4], ah[3], al[3]);
                split_tf32(wr[8 * ks], bh[0], bl[0]);
                split_tf32(wr[8 * ks + 4], bh[1], bl[1]);
                mma_tf32_16x8x8(c, al, bh);
                mma_tf32_16x8x8(c, ah, bl);
                mma_tf32_16x8x8(c, ah, bh);
            }
            const int r0 = 16 * mt + g, j0 = 8 * nt + 2 * t;
            emit(r0, j0, c[0]);
            emit(r0, j0 + 1, c[1]);
            emit(r0 + 8, j0, c[2]);
            emit(r0 + 8, j0 + 1, c[3]);
        }
    } else {
        constexpr int RG = THREADS / DOUT;      // row groups
        constexpr int RPT = TM / RG;            // rows per thread
        static_assert(RG >= 1 && RG <= TM && TM % RG == 0, "unsupported DOUT");
        const int j = tid % DOUT, rg = tid / DOUT;
        float acc[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) acc[k] = 0.f;
#pragma unroll 4
        for (int i = 0; i < DIN; i += 4) {
            const float w0 = sW[(i + 0) * LDW + j], w1 = sW[(i + 1) * LDW + j], w2 = sW[(i + 2) * LDW + j],
                        w3 = sW[(i + 3) * LDW + j];
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const float4 zv = *reinterpret_cast<const float4*>(sZ + (rg + k * RG) * LDZ + i);
                acc[k] = fmaf(zv.x, w0, acc[k]);
                acc[k] = fmaf(zv.y, w1, acc[k]);
                acc[k] = fmaf(zv.z, w2, acc[k]);
                acc[k] = fmaf(zv.w, w3, acc[k]);
            }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k) emit(rg + k * RG, j, acc[k]);
    }
    __syncthreads();

    // phase 3: row norm, write un-normalised output and the normalised concat slice - 128-bit accesses: lane l
    // of the first DOUT / 4 lanes owns columns 4l .. 4l + 3 of the row (shared, global and peer stores alike)
    constexpr int LPO = DOUT / 4;
    for (int rl = warp; rl < TM; rl += WARPS) {
        const int row = row0 + rl;
        if (row >= row_end) continue;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (lane < LPO) v = *reinterpret_cast<const float4*>(sY + rl * LDY + lane * 4);
        float ss = v.x * v.x;
        ss = fmaf(v.y, v.y, ss); ss = fmaf(v.z, v.z, ss); ss = fmaf(v.w, v.w, ss);
        ss = warp_sum(ss);
        const float denom = fmaxf(sqrtf(ss), NORM_EPS);
        if (lane < LPO) {
            const size_t o = static_cast<size_t>(row) * DOUT + lane * 4;
            *reinterpret_cast<float4*>(h_out + o) = v;
            if (PUSH) peer_store(po, o, v);
            if (concat)
                *reinterpret_cast<float4*>(concat + static_cast<size_t>(row) * cstride + out_off + lane * 4) =
                    make_float4(v.x / denom, v.y / denom, v.z / denom, v.w / denom);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// EPI selects the two tile GEMMs of phase 2 (dW2 += dpre^T z and dz = dpre W2): 0 = fp32 FFMA from shared
// memory, 1 = tensor cores (mma.sync m16n8k8 3xTF32, see FwdSmem).  With the MMA the three shared arrays
// get a row stride of width + 8 floats: the fragment loads that walk ROWS with the lane's t index (A of the
// dW product = dpre^T, both B operands) then hit 32 different banks, the one that walks rows with g (A of
// the dz product) is 2-way conflicted.
template <int DIN, int DOUT, int EPI = 0>
struct BwdCfg {
    static constexpr int BJ = (DOUT * DIN >= 4096) ? 4 : ((DOUT * DIN >= 2048) ? 2 : 1);
    static constexpr int NBLOCKS = DOUT * DIN / (4 * BJ);
    static constexpr int NB = (NBLOCKS + THREADS - 1) / THREADS;
    static constexpr int LDP = EPI ? DOUT + 8 : DOUT;       // sDP row stride
    static constexpr int LDZ = EPI ? DIN + 8 : DIN;         // sZ
    static constexpr int LDW = EPI ? DIN + 8 : DIN;         // sW (W2 as given, [DOUT][DIN])
    static constexpr int WT = (DOUT / 16) * (DIN / 8);      // 16 x 8 tiles of dW2
    static constexpr int TPW = WT > 0 ? (WT + WARPS - 1) / WARPS : 1;    // ... per warp
    static constexpr int LDD = DIN + 4;                     // dz tile of the tensor-core path (row-coalesced write-out)
    static constexpr size_t bytes = sizeof(float) * (DOUT * LDW + TM * LDP + TM * LDZ + COOP_FLOATS + (EPI ? TM * LDD : 0)) + 16;
};

// resident CTAs per SM of the persistent grid (measured on B200: 5 beats 4 for the narrow layers - 51 registers
// suffice -, 4 beats both 3 and 5 for 64 x 64, whose accumulators spill below 64 registers)
template <int DIN, int DOUT>
constexpr int bwd_ctas_per_sm() { return DIN * DOUT <= 2048 ? 5 : (DIN * DOUT <= 4096 ? 4 : 1); }

template <int DIN, int DOUT, bool PUSH, int EPI>
__global__ void __launch_bounds__(THREADS, (bwd_ctas_per_sm<DIN, DOUT>()))
layer_bwd_kernel(const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_dst,
                 const float* __restrict__ w_csc, int row_begin, int row_end, int n_tiles,
                 const float* __restrict__ d_concat, const int32_t* __restrict__ dc_slot, int64_t cstride,
                 int64_t out_off,
                 const float* __restrict__ nxt_dx_direct, const float* __restrict__ nxt_dh_nb,
                 const float* __restrict__ h_in, const float* __restrict__ h_nb,
                 const float* __restrict__ act, const float* __restrict__ h_out,
                 const float* __restrict__ drop_mask, const float* __restrict__ W2, float slope,
                 float* __restrict__ dx_direct, float* __restrict__ dh_nb, float* __restrict__ partials,
                 PeerOut po) {
    using C = BwdCfg<DIN, DOUT, EPI>;
    constexpr int BJ = C::BJ, NBLOCKS = C::NBLOCKS, NB = C::NB;
    constexpr int LDP = C::LDP, LDZ = C::LDZ, LDW = C::LDW, WT = C::WT, TPW = C::TPW;
    constexpr int LPI = DIN / 4, LPO = DOUT / 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sDP = reinterpret_cast<float*>(smem_raw);     // [TM][LDP]
    float* sZ = sDP + TM * LDP;                          // [TM][LDZ]
    float* sW = sZ + TM * LDZ;                           // [DOUT][LDW], W2 as given
    float4* sPart = reinterpret_cast<float4*>(sW + DOUT * LDW);
    int* sLong = reinterpret_cast<int*>(sPart + WARPS * 32);
    float* sDZ = reinterpret_cast<float*>(sLong + TM);      // EPI 1: [TM][LDD]  (COOP_FLOATS = 4 * WARPS * 32 + TM floats above)
    constexpr int LDD = C::LDD;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int idx = tid; idx < DOUT * DIN / 4; idx += THREADS)
        *reinterpret_cast<float4*>(sW + (idx / LPI) * LDW + (idx % LPI) * 4) =
            __ldg(reinterpret_cast<const float4*>(W2) + idx);

    float acc[NB][BJ][4];          // EPI 0: dW2 blocks of this thread
#pragma unroll
    for (int n = 0; n < NB; ++n)
#pragma unroll
        for (int a = 0; a < BJ; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[n][a][b] = 0.f;
    float cw[TPW][4];              // EPI 1: dW2 accumulator fragments of this warp's 16 x 8 tiles
#pragma unroll
    for (int q = 0; q < TPW; ++q)
#pragma unroll
        for (int b = 0; b < 4; ++b) cw[q][b] = 0.f;
    const int fg = lane >> 2, ft = lane & 3;      // fragment coordinates of this lane

    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = row_begin + tile * TM;
        __syncthreads();      // previous tile's readers are done with sDP / sZ (also orders the sW fill)

        // phase 1: dpre and z for every row of the tile
        // everything after the transposed aggregation: g = gathered next-layer term of this row
        // `ll`: this lane's index inside the group of lanes that shares the row (the whole warp when the row
        // was gathered by the warp; LPR = max(DIN, DOUT) / 4 lanes when several rows share a warp)
        // `gmask`: the lanes that share the row - the branches below are uniform across them, not across the warp
        auto process_row = [&](int row, int rl, float4 g, int ll, unsigned gmask) {
            float4 dp = make_float4(0.f, 0.f, 0.f, 0.f), z = dp;
            if (row < row_end) {
                const int lo = ll % LPO;
                const size_t oo = static_cast<size_t>(row) * DOUT + lo * 4;
                if (nxt_dx_direct) {
                    const float4 t = ldg4(nxt_dx_direct + oo);
                    g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
                }
                const float4 y = ldg4(h_out + oo);
                const float* dcp = nullptr;         // this row's slice of the concat gradient, if any
                if (dc_slot) {
                    const int slot = __ldg(dc_slot + row);
                    if (slot >= 0) dcp = d_concat + static_cast<size_t>(slot) * cstride + out_off + lo * 4;
                } else if (d_concat) {
                    dcp = d_concat + static_cast<size_t>(row) * cstride + out_off + lo * 4;
                }
                if (dcp) {
                    const float4 dc = ldg4(dcp);
                    float ss = y.x * y.x;
                    ss = fmaf(y.y, y.y, ss); ss = fmaf(y.z, y.z, ss); ss = fmaf(y.w, y.w, ss);
                    ss = group_sum<LPO>(ss, gmask);
                    const float n = sqrtf(ss);
                    if (n >= NORM_EPS) {          // clamp_min passes the gradient to the norm
                        const float4 c = make_float4(y.x / n, y.y / n, y.z / n, y.w / n);
                        float dot = c.x * dc.x;
                        dot = fmaf(c.y, dc.y, dot); dot = fmaf(c.z, dc.z, dot); dot = fmaf(c.w, dc.w, dot);
                        dot = group_sum<LPO>(dot, gmask);
                        g.x += (dc.x - c.x * dot) / n;
                        g.y += (dc.y - c.y * dot) / n;
                        g.z += (dc.z - c.z * dot) / n;
                        g.w += (dc.w - c.w * dot) / n;
                    } else {                      // row clamped to eps: y / eps, constant denominator
                        g.x += dc.x / NORM_EPS; g.y += dc.y / NORM_EPS;
                        g.z += dc.z / NORM_EPS; g.w += dc.w / NORM_EPS;
                    }
                }
                if (drop_mask) {
                    const float4 m = ldg4(drop_mask + oo);
                    g.x *= m.x; g.y *= m.y; g.z *= m.z; g.w *= m.w;
                }
                const float4 a = act ? ldg4(act + oo) : y;
                dp.x = a.x > 0.f ? g.x : g.x * slope;
                dp.y = a.y > 0.f ? g.y : g.y * slope;
                dp.z = a.z > 0.f ? g.z : g.z * slope;
                dp.w = a.w > 0.f ? g.w : g.w * slope;
                const size_t oi = static_cast<size_t>(row) * DIN + (ll % LPI) * 4;
                const float4 x = ldg4(h_in + oi), hn = ldg4(h_nb + oi);
                z = make_float4(x.x * hn.x, x.y * hn.y, x.z * hn.z, x.w * hn.w);
            }
            if (ll < LPO) *reinterpret_cast<float4*>(sDP + rl * LDP + ll * 4) = dp;
            if (ll < LPI) *reinterpret_cast<float4*>(sZ + rl * LDZ + ll * 4) = z;
        };
        if (!nxt_dh_nb) {
            // dense-only call (the transposed aggregation ran as csc_gather in front, or this is the last
            // layer): no warp-wide gather, so a row needs only LPR = max(DIN, DOUT) / 4 lanes and 32 / LPR rows
            // share a warp - twice (width 64) to four times (width 32) the loads in flight per warp and no
            // lanes that merely duplicate their neighbours' work
            constexpr int LPR = (LPI > LPO ? LPI : LPO);
            constexpr int RPW = 32 / LPR;
            const int sub = lane / LPR, ll = lane % LPR;
            const unsigned gmask = (LPR == 32 ? FULL_MASK : ((1u << LPR) - 1u)) << (sub * LPR);
            for (int rl = warp * RPW + sub; rl < TM; rl += WARPS * RPW)
                process_row(row0 + rl, rl, make_float4(0.f, 0.f, 0.f, 0.f), ll, gmask);
            __syncthreads();
        } else {
        bool warp_long = false;
        for (int rl = warp; rl < TM; rl += WARPS) {
            const int row = row0 + rl;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            bool is_long = false;
            if (row < row_end) {
                const int beg = __ldg(csc_ptr + row), end = __ldg(csc_ptr + row + 1);
                if (end - beg > LONG_ROW) is_long = true;
                else g = gather_row<DOUT>(csc_dst, w_csc, beg, end, nxt_dh_nb, lane);
            }
            if (!is_long) process_row(row, rl, g, lane, FULL_MASK);
            if (lane == 0) sLong[rl] = is_long ? 1 : 0;
            warp_long |= is_long;
        }
        // the barrier that ends phase 1 also tells every thread whether the tile holds a long row
        if (__syncthreads_or(warp_long ? 1 : 0)) {
            for (int rl = 0; rl < TM; ++rl) {
                if (!sLong[rl]) continue;                 // CTA-uniform
                const int row = row0 + rl;
                const int beg = __ldg(csc_ptr + row), end = __ldg(csc_ptr + row + 1);
                const float4 g = gather_row_cta<DOUT>(csc_dst, w_csc, beg, end, nxt_dh_nb, sPart, warp, lane);
                if (warp == rl % WARPS) process_row(row, rl, g, lane, FULL_MASK);
            }
            __syncthreads();
        }
        }

        if constexpr (EPI == 1) {
            static_assert(TM == 32 && DIN % 8 == 0 && DOUT % 16 == 0, "m16n8k8 tiling");
            constexpr int NTN = DIN / 8;
            // phase 2a: dW2 (DOUT x DIN) += dpre^T (DOUT x 32) z (32 x DIN): A[m][k] = sDP[k][m], B[k][n] = sZ[k][n]
#pragma unroll
            for (int ks = 0; ks < TM / 8; ++ks) {
#pragma unroll
                for (int q = 0; q < TPW; ++q) {
                    const int p = warp + q * WARPS;
                    if (p < WT) {                                     // warp-uniform
                        const int mt = p / NTN, nt = p % NTN;
                        const float* ap = sDP + (8 * ks + ft) * LDP + 16 * mt + fg;
                        const float* bp = sZ + (8 * ks + ft) * LDZ + 8 * nt + fg;
                        uint32_t ah[4], al[4], bh[2], bl[2];
                        split_tf32(ap[0], ah[0], al[0]);
                        split_tf32(ap[8], ah[1], al[1]);
                        split_tf32(ap[4 * LDP], ah[2], al[2]);
                        split_tf32(ap[4 * LDP + 8], ah[3], al[3]);
                        split_tf32(bp[0], bh[0], bl[0]);
                        split_tf32(bp[4 * LDZ], bh[1], bl[1]);
                        mma_tf32_16x8x8(cw[q], al, bh);
                        mma_tf32_16x8x8(cw[q], ah, bl);
                        mma_tf32_16x8x8(cw[q], ah, bh);
                    }
                }
            }
            // phase 2b: dz (32 x DIN) = dpre (32 x DOUT) W2 (DOUT x DIN): A[m][k] = sDP[m][k], B[k][n] = sW[k][n]
            for (int p = warp; p < 2 * NTN; p += WARPS) {
                const int mt = p / NTN, nt = p % NTN;
                const float* a0p = sDP + (16 * mt + fg) * LDP + ft;
                const float* a1p = a0p + 8 * LDP;
                const float* bp = sW + ft * LDW + 8 * nt + fg;
                float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < DOUT / 8; ++ks) {
                    uint32_t ah[4], al[4], bh[2], bl[2];
                    split_tf32(a0p[8 * ks], ah[0], al[0]);
                    split_tf32(a1p[8 * ks], ah[1], al[1]);
                    split_tf32(a0p[8 * ks + 4], ah[2], al[2]);
                    split_tf32(a1p[8 * ks + 4], ah[3], al[3]);
                    split_tf32(bp[(8 * ks) * LDW], bh[0], bl[0]);
                    split_tf32(bp[(8 * ks + 4) * LDW], bh[1], bl[1]);
                    mma_tf32_16x8x8(c, al, bh);
                    mma_tf32_16x8x8(c, ah, bl);
                    mma_tf32_16x8x8(c, ah, bh);
                }
                // dz fragment -> shared memory; the write-out below is row-coalesced
                *reinterpret_cast<float2*>(sDZ + (16 * mt + fg) * LDD + 8 * nt + 2 * ft) = make_float2(c[0], c[1]);
                *reinterpret_cast<float2*>(sDZ + (16 * mt + fg + 8) * LDD + 8 * nt + 2 * ft) = make_float2(c[2], c[3]);
            }
            __syncthreads();
            // phase 2c: dx_direct = dz (.) h_nb, dh_nb = dz (.) h_in, whole rows per lane group: every global (and
            // peer) store instruction covers complete 128-byte lines of a row (the fragment layout stored 32-byte
            // pieces of 16 rows - over NVLink that is one packet per 32 bytes; measured at 4 GPUs: the dense kernel
            // of a rank took LONGER than the single-GPU one although it owned a quarter of the rows)
            {
                constexpr int RPW2 = 32 / LPI;
                const int sub = lane / LPI, ll = lane % LPI;
                for (int rl = warp * RPW2 + sub; rl < TM; rl += WARPS * RPW2) {
                    const int row = row0 + rl;
                    if (row < row_end) {
                        const float4 d = *reinterpret_cast<const float4*>(sDZ + rl * LDD + ll * 4);
                        const size_t o = static_cast<size_t>(row) * DIN + ll * 4;
                        const float4 hn = ldg4(h_nb + o), hi = ldg4(h_in + o);
                        *reinterpret_cast<float4*>(dx_direct + o) = make_float4(d.x * hn.x, d.y * hn.y, d.z * hn.z, d.w * hn.w);
                        const float4 dh = make_float4(d.x * hi.x, d.y * hi.y, d.z * hi.z, d.w * hi.w);
                        *reinterpret_cast<float4*>(dh_nb + o) = dh;
                        if (PUSH) peer_store(po, o, dh);
                    }
                }
            }
        } else {
        // phase 2a: dW2 += dpre^T z
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int b = tid + n * THREADS;
            if (b < NBLOCKS) {
                const int i0 = (b % LPI) * 4, j0 = (b / LPI) * BJ;
#pragma unroll 4
                for (int r = 0; r < TM; ++r) {
                    const float4 zv = *reinterpret_cast<const float4*>(sZ + r * LDZ + i0);
#pragma unroll
                    for (int a = 0; a < BJ; ++a) {
                        const float d = sDP[r * LDP + j0 + a];
                        acc[n][a][0] = fmaf(d, zv.x, acc[n][a][0]);
                        acc[n][a][1] = fmaf(d, zv.y, acc[n][a][1]);
                        acc[n][a][2] = fmaf(d, zv.z, acc[n][a][2]);
                        acc[n][a][3] = fmaf(d, zv.w, acc[n][a][3]);
                    }
                }
            }
        }

        // phase 2b: dz = dpre W2; dx_direct = dz (.) h_nb; dh_nb = dz (.) h_in
        {
            constexpr int RG = THREADS / DIN;
            constexpr int RPT = TM / RG;
            static_assert(RG >= 1 && RG <= TM && TM % RG == 0, "unsupported DIN");
            const int i = tid % DIN, rg = tid / DIN;
            float dz[RPT];
#pragma unroll
            for (int k = 0; k < RPT; ++k) dz[k] = 0.f;
#pragma unroll 4
            for (int j = 0; j < DOUT; j += 4) {
                const float w0 = sW[(j + 0) * LDW + i], w1 = sW[(j + 1) * LDW + i], w2 = sW[(j + 2) * LDW + i],
                            w3 = sW[(j + 3) * LDW + i];
#pragma unroll
                for (int k = 0; k < RPT; ++k) {
                    const float4 d = *reinterpret_cast<const float4*>(sDP + (rg + k * RG) * LDP + j);
                    dz[k] = fmaf(d.x, w0, dz[k]);
                    dz[k] = fmaf(d.y, w1, dz[k]);
                    dz[k] = fmaf(d.z, w2, dz[k]);
                    dz[k] = fmaf(d.w, w3, dz[k]);
                }
            }
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const int row = row0 + rg + k * RG;
                if (row < row_end) {
                    const size_t o = static_cast<size_t>(row) * DIN + i;
                    dx_direct[o] = dz[k] * __ldg(h_nb + o);
                    const float dh = dz[k] * __ldg(h_in + o);
                    dh_nb[o] = dh;
                    if (PUSH) peer_store(po, o, dh);
                }
            }
        }
        }   // EPI == 0
    }

    // one dW2 partial per CTA
    float* my = partials + static_cast<size_t>(blockIdx.x) * DOUT * DIN;
    if constexpr (EPI == 1) {
#pragma unroll
        for (int q = 0; q < TPW; ++q) {
            const int p = warp + q * WARPS;
            if (p < WT) {
                const int mt = p / (DIN / 8), nt = p % (DIN / 8);
                float* o = my + (16 * mt + fg) * DIN + 8 * nt + 2 * ft;
                *reinterpret_cast<float2*>(o) = make_float2(cw[q][0], cw[q][1]);
                *reinterpret_cast<float2*>(o + 8 * DIN) = make_float2(cw[q][2], cw[q][3]);
            }
        }
    } else {
#pragma unroll
        for (int n = 0; n < NB; ++n) {
            const int b = tid + n * THREADS;
            if (b < NBLOCKS) {
                const int i0 = (b % LPI) * 4, j0 = (b / LPI) * BJ;
#pragma unroll
                for (int a = 0; a < BJ; ++a)
                    *reinterpret_cast<float4*>(my + (j0 + a) * DIN + i0) =
                        make_float4(acc[n][a][0], acc[n][a][1], acc[n][a][2], acc[n][a][3]);
            }
        }
    }
}

// dW2[idx] = sum_p partials[p][idx], fixed order, accumulated in fp64 (the per-CTA partials are
// fp32 sums over a few hundred rows; the cross-CTA sum must not add its own rounding on top).
// One CTA per 32 consecutive elements (one 128-byte line of every partial): lane = (partial slot
// lane / 8, float4 lane % 8), warp w and slot g take partials p = 4 w + g (mod 32), so every load
// instruction reads four full lines and a thread walks n_partials / 32 of them; slots meet by shuffle,
// warps in shared memory, always in the same order.
__global__ void __launch_bounds__(256)
reduce_partials_kernel(const float* __restrict__ partials, int n_partials, int n_elem, float* __restrict__ out) {
    __shared__ double sAcc[8][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e4 = lane & 7, slot = lane >> 3;
    const int idx = blockIdx.x * 32 + e4 * 4;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    if (idx < n_elem) {
#pragma unroll 4
        for (int p = warp * 4 + slot; p < n_partials; p += 32) {
            const float4 v = ldg4(partials + static_cast<size_t>(p) * n_elem + idx);
            s0 += static_cast<double>(v.x); s1 += static_cast<double>(v.y);
            s2 += static_cast<double>(v.z); s3 += static_cast<double>(v.w);
        }
    }
#pragma unroll
    for (int o = 8; o < 32; o <<= 1) {
        s0 += __shfl_xor_sync(FULL_MASK, s0, o); s1 += __shfl_xor_sync(FULL_MASK, s1, o);
        s2 += __shfl_xor_sync(FULL_MASK, s2, o); s3 += __shfl_xor_sync(FULL_MASK, s3, o);
    }
    if (slot == 0) {
        sAcc[warp][e4 * 4 + 0] = s0; sAcc[warp][e4 * 4 + 1] = s1;
        sAcc[warp][e4 * 4 + 2] = s2; sAcc[warp][e4 * 4 + 3] = s3;
    }
    __syncthreads();
    if (threadIdx.x < 32 && blockIdx.x * 32 + threadIdx.x < n_elem) {
        double t = sAcc[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < 8; ++w) t += sAcc[w][threadIdx.x];
        out[blockIdx.x * 32 + threadIdx.x] = static_cast<float>(t);
    }
}

// Transposed aggregation of one block of WARPS rows: warp w gathers row (first + w) over its CSC segment.
// A row with more than LONG_ROW edges (a hot entity of a power-law KG) is gathered by ALL warps of the CTA
// (gather_row_cta: fixed slices, partials added in warp order - deterministic) before the short rows, so no
// single warp walks 10^4 edges while its CTA idles.  Returns this warp's row total (every lane group holds it).
template <int D>
__device__ __forceinline__ float4 gather_block_rows(const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_dst,
                                                    const float* __restrict__ w_csc, const float* __restrict__ feat,
                                                    int first, int row_end, float4* sPart, int warp, int lane,
                                                    const float* pre) {
    const int row = first + warp;
    int beg = 0, end = 0;
    if (row < row_end) { beg = __ldg(csc_ptr + row); end = __ldg(csc_ptr + row + 1); }
    const int long_row = pre ? SPLIT_ROW : LONG_ROW;           // with the pre-pass every row above SPLIT_ROW is done
    const bool is_long = end - beg > long_row;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (__syncthreads_or(is_long ? 1 : 0)) {
        for (int w = 0; w < WARPS; ++w) {                       // CTA-uniform: every thread re-reads the bounds
            const int r = first + w;
            if (r >= row_end) break;
            const int b = __ldg(csc_ptr + r), e = __ldg(csc_ptr + r + 1);
            if (e - b <= long_row) continue;
            if (pre && e - b > SPLIT_ROW) {      // gathered by the pre-pass (split_prepass) into row r of `pre`
                if (w == warp) g = *reinterpret_cast<const float4*>(pre + static_cast<size_t>(r) * D + (lane % (D / 4)) * 4);
                continue;
            }
            const float4 t = gather_row_cta<D>(csc_dst, w_csc, b, e, feat, sPart, warp, lane);
            if (w == warp) g = t;
        }
    }
    if (!is_long && row < row_end) g = gather_row<D>(csc_dst, w_csc, beg, end, feat, lane);
    return g;
}

// out[v] = (add ? add[v] : 0) + sum_{k in CSC row v} w[k] * feat[dst[k]]: the gradient that reaches the output
// of a layer through the NEXT layer (its direct term plus the transposed aggregation of its dh_nb).  A kernel
// of its own in front of layer_bwd: as phase 1 of that kernel the gather ran at 4 CTAs per SM (64 registers,
// persistent dW2 accumulators) in lockstep with the tile GEMMs - 0.235 ms for the 64 x 64 layer where this
// kernel plus the dense remainder take less (measured, profiles/README.md).
template <int D>
__global__ void __launch_bounds__(THREADS, 8)      // 32 registers: 8 CTAs per SM (at 40 the kernel ran 10 % slower)
csc_gather_kernel(const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_dst,
                  const float* __restrict__ w_csc, int row_begin, int row_end, const float* __restrict__ add,
                  const float* __restrict__ feat, float* __restrict__ out, int split_done) {
    constexpr int LP = D / 4;
    __shared__ float4 sPart[WARPS * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int first = row_begin + blockIdx.x * WARPS, row = first + warp;
    float4 g = gather_block_rows<D>(csc_ptr, csc_dst, w_csc, feat, first, row_end, sPart, warp, lane,
                                    split_done ? out : nullptr);
    if (row < row_end && lane < LP) {
        const size_t o = static_cast<size_t>(row) * D + lane * 4;
        if (add) {
            const float4 t = ldg4(add + o);
            g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        }
        *reinterpret_cast<float4*>(out + o) = g;
    }
}

// d_entity[v] = d_concat[v, in_off:+D] + dx_direct[v] + sum_{k in CSC row v} w[k] * dh_nb[dst[k]]
template <int D, bool PUSH>
__global__ void __launch_bounds__(THREADS, 8)
input_bwd_kernel(const int32_t* __restrict__ csc_ptr, const int32_t* __restrict__ csc_dst,
                 const float* __restrict__ w_csc, int row_begin, int row_end,
                 const float* __restrict__ d_concat, const int32_t* __restrict__ dc_slot, int64_t cstride,
                 int64_t in_off, const float* __restrict__ dx_direct, const float* __restrict__ dh_nb,
                 float* __restrict__ d_entity, int split_done, PeerOut po) {
    constexpr int LP = D / 4;
    __shared__ float4 sPart[WARPS * 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int first = row_begin + blockIdx.x * WARPS, row = first + warp;
    float4 g = gather_block_rows<D>(csc_ptr, csc_dst, w_csc, dh_nb, first, row_end, sPart, warp, lane,
                                    split_done ? d_entity : nullptr);
    if (row < row_end && lane < LP) {
        const size_t o = static_cast<size_t>(row) * D + lane * 4;
        const float4 t = ldg4(dx_direct + o);
        g.x += t.x; g.y += t.y; g.z += t.z; g.w += t.w;
        const float* dcp = nullptr;
        if (dc_slot) {
            const int slot = __ldg(dc_slot + row);
            if (slot >= 0) dcp = d_concat + static_cast<size_t>(slot) * cstride + in_off + lane * 4;
        } else if (d_concat) {
            dcp = d_concat + static_cast<size_t>(row) * cstride + in_off + lane * 4;
        }
        if (dcp) {
            const float4 c = ldg4(dcp);
            g.x += c.x; g.y += c.y; g.z += c.z; g.w += c.w;
        }
        *reinterpret_cast<float4*>(d_entity + o) = g;
        if (PUSH) peer_store(po, o, g);
    }
}

// ------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------
struct FwdArgs {
    const kgat_graph_t* g; int row_begin, row_end;
    const float *h_in, *w_csr, *W2, *drop_mask; float slope;
    float *h_nb, *act, *h_out, *concat; int64_t cstride, out_off, in_off; cudaStream_t st;
    bool push; PeerOut po;
};

// Tile GEMM of layer_fwd: tensor cores (mma.sync 3xTF32) for input widths >= 32, fp32 FFMA below.  The 3xTF32
// products carry ~7e-7 of relative error instead of fp32's 6e-8; the embeddings stay far inside the 1e-5
// bar at every width, but at width 16 the extra sign flips of near-zero pre-activations (LeakyReLU) showed
// up in the entity gradient (2.9e-5 of its largest entry against a bound of twice the fp32 reference's own
// 0.8e-5), so the narrow layers keep FFMA.  KGAT_FWD_EPILOGUE=ffma / mma forces one of them (A/B runs).
inline int fwd_epilogue(int d_in) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = std::getenv("KGAT_FWD_EPILOGUE");
        forced = !e ? -1 : (std::strcmp(e, "mma") == 0 ? 1 : (std::strcmp(e, "ffma") == 0 ? 0 : -1));
    }
    return forced >= 0 ? forced : (d_in >= 32 ? 1 : 0);
}

template <int DIN, int DOUT, bool PUSH, int EPI>
int launch_fwd_t(const FwdArgs& a) {
    constexpr size_t smem = FwdSmem<DIN, DOUT, EPI>::bytes;
    static PerDeviceOnce configured;
    if (configured.need()) {
        KGAT_CUDA(cudaFuncSetAttribute(layer_fwd_kernel<DIN, DOUT, PUSH, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        configured.set();
    }
    const int rows = a.row_end - a.row_begin;
    const unsigned grid = (rows + TM - 1) / TM;
    bool split_done = false;
    const int rc = split_prepass<DIN>(a.g, 0, a.w_csr, a.h_in, a.row_begin, a.row_end, a.h_nb, a.st, &split_done);
    if (rc != KGAT_OK) return rc;
    layer_fwd_kernel<DIN, DOUT, PUSH, EPI><<<grid, THREADS, smem, a.st>>>(
        a.g->csr_ptr, a.g->csr_src, a.w_csr, a.row_begin, a.row_end, a.h_in, a.W2, a.drop_mask, a.slope,
        a.h_nb, a.act, a.h_out, a.concat, a.cstride, a.out_off, a.in_off, split_done ? 1 : 0, a.po);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

template <int DIN, int DOUT>
int launch_fwd(const FwdArgs& a) {
    if (fwd_epilogue(DIN) == 1) return a.push ? launch_fwd_t<DIN, DOUT, true, 1>(a) : launch_fwd_t<DIN, DOUT, false, 1>(a);
    return a.push ? launch_fwd_t<DIN, DOUT, true, 0>(a) : launch_fwd_t<DIN, DOUT, false, 0>(a);
}

struct BwdArgs {
    const kgat_graph_t* g; int row_begin, row_end;
    const float* d_concat; const int32_t* dc_slot; int64_t cstride, out_off;
    const float *nxt_dx_direct, *nxt_dh_nb, *w_csc, *h_in, *h_nb, *act, *h_out, *drop_mask, *W2; float slope;
    float *dx_direct, *dh_nb, *dW2, *partials; cudaStream_t st;
    bool push; PeerOut po;
};

inline int bwd_max_ctas() { return 5 * num_sms(); }      // upper bound over the instantiations: sizes the partials

// Tile GEMMs of layer_bwd: tensor cores (widths DIN >= 32, DOUT >= 16), single-GPU and partitioned (PUSH)
// instantiations alike, so one GPU and N GPUs run the same arithmetic (tests/test_gpu_dist.py compares them
// directly).  KGAT_BWD_EPILOGUE=mma / ffma forces one.
inline int bwd_epilogue(int d_in, int d_out, bool push) {
    static int forced = -2;
    if (forced == -2) {
        const char* e = std::getenv("KGAT_BWD_EPILOGUE");
        forced = !e ? -1 : (std::strcmp(e, "mma") == 0 ? 1 : (std::strcmp(e, "ffma") == 0 ? 0 : -1));
    }
    (void)push;
    const int want = forced >= 0 ? forced : 1;
    return (want == 1 && d_in >= 32 && d_out >= 16) ? 1 : 0;
}

// CTAs (= dW2 partials) of the persistent layer_bwd grid over `rows` rows
template <int DIN, int DOUT>
int bwd_grid(int rows) {
    const int n_tiles = (rows + TM - 1) / TM;
    const int cap = bwd_ctas_per_sm<DIN, DOUT>() * num_sms();
    return n_tiles < cap ? n_tiles : cap;
}

template <int DIN, int DOUT, bool PUSH, int EPI>
int launch_bwd_t(const BwdArgs& a) {
    constexpr size_t smem = BwdCfg<DIN, DOUT, EPI>::bytes;
    static PerDeviceOnce configured;
    if (configured.need()) {
        KGAT_CUDA(cudaFuncSetAttribute(layer_bwd_kernel<DIN, DOUT, PUSH, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
        configured.set();
    }
    const int rows = a.row_end - a.row_begin;
    const int n_tiles = (rows + TM - 1) / TM;
    const int grid = bwd_grid<DIN, DOUT>(rows);
    layer_bwd_kernel<DIN, DOUT, PUSH, EPI><<<grid, THREADS, smem, a.st>>>(
        a.g->csc_ptr, a.g->csc_dst, a.w_csc, a.row_begin, a.row_end, n_tiles, a.d_concat, a.dc_slot, a.cstride,
        a.out_off,
        a.nxt_dx_direct, a.nxt_dh_nb, a.h_in, a.h_nb, a.act, a.h_out, a.drop_mask, a.W2, a.slope, a.dx_direct,
        a.dh_nb, a.partials, a.po);
    KGAT_LAUNCH_CHECK();
    if (!a.dW2) return KGAT_OK;          // the caller reduces the partials itself (kgat_layer_bwd_reduce)
    const int n_elem = DIN * DOUT;
    reduce_partials_kernel<<<(n_elem + 31) / 32, 256, 0, a.st>>>(a.partials, grid, n_elem, a.dW2);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

template <int DIN, int DOUT>
int launch_bwd(const BwdArgs& a) {
    if constexpr (DIN >= 32 && DOUT >= 16) {
        if (bwd_epilogue(DIN, DOUT, a.push) == 1)
            return a.push ? launch_bwd_t<DIN, DOUT, true, 1>(a) : launch_bwd_t<DIN, DOUT, false, 1>(a);
    }
    return a.push ? launch_bwd_t<DIN, DOUT, true, 0>(a) : launch_bwd_t<DIN, DOUT, false, 0>(a);
}

#define KGAT_DIM_PAIRS(X) \
    X(128, 128) X(128, 64) X(64, 64) X(64, 32) X(32, 32) X(32, 16) X(16, 16) X(16, 8)

}  // namespace
}  // namespace kgat

using namespace kgat;

// pg == nullptr: the single-GPU entry points; otherwise the exchanged output must lie in the own arena
extern "C" int kgat_layer_fwd_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                                   const float* h_in, const float* w_csr, const float* W2,
                                   const float* drop_mask, int d_in, int d_out, float slope,
                                   float* h_nb, float* act, float* h_out, float* concat,
                                   int64_t concat_stride, int64_t out_off, int64_t in_off,
                                   const kgat_peer_group_t* pg, kgat_stream_t stream) {
    KGAT_REQUIRE(g && h_in && W2 && h_nb && h_out);
    KGAT_REQUIRE(g->n_edges == 0 || w_csr);
    KGAT_REQUIRE(row_begin >= 0 && row_end <= g->n_nodes && row_begin <= row_end);
    KGAT_REQUIRE(aligned16(h_in) && aligned16(h_nb) && aligned16(h_out) && aligned16(W2));
    KGAT_REQUIRE(!drop_mask || act);        // with dropout the pre-mask activation must be kept
    if (concat) {
        KGAT_REQUIRE(aligned16(concat) && concat_stride % 4 == 0 && out_off % 4 == 0 && out_off >= 0);
        KGAT_REQUIRE(in_off < 0 || in_off % 4 == 0);
    } else {
        in_off = -1;
    }
    PeerOut po = no_peers();
    if (pg) KGAT_REQUIRE(d_out > 0 && peer_aliases(pg, h_out, sizeof(float) * g->n_nodes * d_out, &po));
    if (row_begin == row_end) return KGAT_OK;
    FwdArgs a{g, (int)row_begin, (int)row_end, h_in, w_csr, W2, drop_mask, slope, h_nb, act, h_out, concat,
              concat_stride, out_off, in_off, as_stream(stream), pg != nullptr, po};
#define X(DI, DO) if (d_in == DI && d_out == DO) return launch_fwd<DI, DO>(a);
    KGAT_DIM_PAIRS(X)
#undef X
    return KGAT_ERR_UNSUPPORTED_DIM;
}

extern "C" int kgat_layer_fwd(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                              const float* h_in, const float* w_csr, const float* W2,
                              const float* drop_mask, int d_in, int d_out, float slope,
                              float* h_nb, float* act, float* h_out, float* concat,
                              int64_t concat_stride, int64_t out_off, int64_t in_off,
                              kgat_stream_t stream) {
    return kgat_layer_fwd_push(g, row_begin, row_end, h_in, w_csr, W2, drop_mask, d_in, d_out, slope, h_nb, act, h_out,
                               concat, concat_stride, out_off, in_off, nullptr, stream);
}

extern "C" int kgat_layer_bwd_partials(void) { return bwd_max_ctas(); }

extern "C" int kgat_layer_bwd_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                                   const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride,
                                   int64_t out_off, const float* nxt_dx_direct, const float* nxt_dh_nb,
                                   const float* w_csc, const float* h_in, const float* h_nb, const float* act,
                                   const float* h_out, const float* drop_mask, const float* W2, int d_in, int d_out,
                                   float slope, float* dx_direct, float* dh_nb, float* dW2, float* dW2_partials,
                                   const kgat_peer_group_t* pg, kgat_stream_t stream) {
    KGAT_REQUIRE(g && h_in && h_nb && h_out && W2 && dx_direct && dh_nb && dW2_partials);
    KGAT_REQUIRE(row_begin >= 0 && row_end <= g->n_nodes && row_begin <= row_end);
    KGAT_REQUIRE(!nxt_dh_nb || g->n_edges == 0 || w_csc);
    KGAT_REQUIRE(!drop_mask || act);
    KGAT_REQUIRE(aligned16(h_in) && aligned16(h_nb) && aligned16(h_out) && aligned16(W2) && aligned16(dx_direct) &&
                 aligned16(dh_nb) && aligned16(dW2_partials));
    if (d_concat) KGAT_REQUIRE(aligned16(d_concat) && concat_stride % 4 == 0 && out_off % 4 == 0 && out_off >= 0);
    PeerOut po = no_peers();
    if (pg) KGAT_REQUIRE(d_in > 0 && peer_aliases(pg, dh_nb, sizeof(float) * g->n_nodes * d_in, &po));
    cudaStream_t st = as_stream(stream);
    if (row_begin == row_end) {
        if (dW2) KGAT_CUDA(cudaMemsetAsync(dW2, 0, sizeof(float) * d_in * d_out, st));
        return KGAT_OK;
    }
    KGAT_REQUIRE(!dc_row_slot || d_concat);
    BwdArgs a{g, (int)row_begin, (int)row_end, d_concat, dc_row_slot, concat_stride, out_off, nxt_dx_direct, nxt_dh_nb, w_csc,
              h_in, h_nb, act, h_out, drop_mask, W2, slope, dx_direct, dh_nb, dW2, dW2_partials, st, pg != nullptr, po};
#define X(DI, DO) if (d_in == DI && d_out == DO) return launch_bwd<DI, DO>(a);
    KGAT_DIM_PAIRS(X)
#undef X
    return KGAT_ERR_UNSUPPORTED_DIM;
}

extern "C" int kgat_layer_bwd(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                              const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride,
                              int64_t out_off, const float* nxt_dx_direct, const float* nxt_dh_nb, const float* w_csc,
                              const float* h_in, const float* h_nb, const float* act, const float* h_out,
                              const float* drop_mask, const float* W2, int d_in, int d_out, float slope,
                              float* dx_direct, float* dh_nb, float* dW2, float* dW2_partials,
                              kgat_stream_t stream) {
    return kgat_layer_bwd_push(g, row_begin, row_end, d_concat, dc_row_slot, concat_stride, out_off, nxt_dx_direct,
                               nxt_dh_nb, w_csc, h_in, h_nb, act, h_out, drop_mask, W2, d_in, d_out, slope, dx_direct,
                               dh_nb, dW2, dW2_partials, nullptr, stream);
}

extern "C" int kgat_layer_bwd_reduce(int64_t rows, int d_in, int d_out, const float* dW2_partials, float* dW2,
                                     kgat_stream_t stream) {
    KGAT_REQUIRE(rows >= 0 && rows < (int64_t(1) << 31) && dW2_partials && dW2 && aligned16(dW2_partials));
    cudaStream_t st = as_stream(stream);
    if (rows == 0) {
        KGAT_CUDA(cudaMemsetAsync(dW2, 0, sizeof(float) * d_in * d_out, st));
        return KGAT_OK;
    }
    int grid = -1;
#define X(DI, DO) if (d_in == DI && d_out == DO) grid = bwd_grid<DI, DO>(static_cast<int>(rows));
    KGAT_DIM_PAIRS(X)
#undef X
    if (grid < 0) return KGAT_ERR_UNSUPPORTED_DIM;
    const int n_elem = d_in * d_out;
    reduce_partials_kernel<<<(n_elem + 31) / 32, 256, 0, st>>>(dW2_partials, grid, n_elem, dW2);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_input_bwd_push(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                                   const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride,
                                   int64_t in_off, const float* dx_direct, const float* dh_nb, const float* w_csc,
                                   int d, float* d_entity, const kgat_peer_group_t* pg, kgat_stream_t stream) {
    KGAT_REQUIRE(g && dx_direct && dh_nb && d_entity);
    KGAT_REQUIRE(g->n_edges == 0 || w_csc);
    KGAT_REQUIRE(row_begin >= 0 && row_end <= g->n_nodes && row_begin <= row_end);
    KGAT_REQUIRE(aligned16(dx_direct) && aligned16(dh_nb) && aligned16(d_entity));
    if (d_concat) KGAT_REQUIRE(aligned16(d_concat) && concat_stride % 4 == 0 && in_off % 4 == 0 && in_off >= 0);
    PeerOut po = no_peers();
    if (pg) KGAT_REQUIRE(d > 0 && peer_aliases(pg, d_entity, sizeof(float) * g->n_nodes * d, &po));
    if (row_begin == row_end) return KGAT_OK;
    const int rows = static_cast<int>(row_end - row_begin);
    const unsigned grid = (rows + WARPS - 1) / WARPS;
    cudaStream_t st = as_stream(stream);
    bool split_done = false;
#define LAUNCH_T(D, P)                                                                                     \
    input_bwd_kernel<D, P><<<grid, THREADS, 0, st>>>(g->csc_ptr, g->csc_dst, w_csc, (int)row_begin,        \
                                                     (int)row_end, d_concat, dc_row_slot, concat_stride,   \
                                                     in_off, dx_direct, dh_nb, d_entity, split_done ? 1 : 0, po)
#define LAUNCH(D)                                                                                          \
    if (d == D) {                                                                                          \
        const int rc = split_prepass<D>(g, 1, w_csc, dh_nb, (int)row_begin, (int)row_end, d_entity, st,    \
                                        &split_done);                                                      \
        if (rc != KGAT_OK) return rc;                                                                      \
        if (pg) LAUNCH_T(D, true); else LAUNCH_T(D, false);                                                \
        KGAT_LAUNCH_CHECK();                                                                               \
        return KGAT_OK;                                                                                    \
    }
    LAUNCH(128) LAUNCH(64) LAUNCH(32) LAUNCH(16) LAUNCH(8)
#undef LAUNCH
#undef LAUNCH_T
    return KGAT_ERR_UNSUPPORTED_DIM;
}

extern "C" int kgat_input_bwd(const kgat_graph_t* g, int64_t row_begin, int64_t row_end,
                              const float* d_concat, const int32_t* dc_row_slot, int64_t concat_stride,
                              int64_t in_off, const float* dx_direct, const float* dh_nb, const float* w_csc, int d,
                              float* d_entity, kgat_stream_t stream) {
    return kgat_input_bwd_push(g, row_begin, row_end, d_concat, dc_row_slot, concat_stride, in_off, dx_direct, dh_nb,
                               w_csc, d, d_entity, nullptr, stream);
}

extern "C" int kgat_csc_gather(const kgat_graph_t* g, int64_t row_begin, int64_t row_end, const float* add,
                               const float* feat, const float* w_csc, int d, float* out, kgat_stream_t stream) {
    KGAT_REQUIRE(g && feat && out);
    KGAT_REQUIRE(g->n_edges == 0 || w_csc);
    KGAT_REQUIRE(row_begin >= 0 && row_end <= g->n_nodes && row_begin <= row_end);
    KGAT_REQUIRE(aligned16(feat) && aligned16(out) && (!add || aligned16(add)));
    if (row_begin == row_end) return KGAT_OK;
    const int rows = static_cast<int>(row_end - row_begin);
    const unsigned grid = (rows + WARPS - 1) / WARPS;
    cudaStream_t st = as_stream(stream);
    bool split_done = false;
#define LAUNCH(D)                                                                                          \
    if (d == D) {                                                                                          \
        const int rc = split_prepass<D>(g, 1, w_csc, feat, (int)row_begin, (int)row_end, out, st,          \
                                        &split_done);                                                      \
        if (rc != KGAT_OK) return rc;                                                                      \
        csc_gather_kernel<D><<<grid, THREADS, 0, st>>>(g->csc_ptr, g->csc_dst, w_csc, (int)row_begin,      \
                                                       (int)row_end, add, feat, out, split_done ? 1 : 0);  \
        KGAT_LAUNCH_CHECK();                                                                               \
        return KGAT_OK;                                                                                    \
    }
    LAUNCH(128) LAUNCH(64) LAUNCH(32) LAUNCH(16) LAUNCH(8)
#undef LAUNCH
    return KGAT_ERR_UNSUPPORTED_DIM;
}
