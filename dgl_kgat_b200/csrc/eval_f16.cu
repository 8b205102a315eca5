// (5) Full-rank user x item scoring on the tensor cores, split-fp16 operands prepared once per call and
// streamed by bulk asynchronous copies, with the per-row top-K fused into the accumulator read-out.
//
// Replaces the per-user loop of the reference (metric.py:44-59): item-matrix re-gather, GEMV, mask, full
// sort, device->host sync per user.  score[u, i] = <emb[user u], emb[item i]> is a (U x dim) x (dim x I)
// GEMM (469 GFLOP for the 53 469 eval users of the Amazon-book shape); the ranking needs fp32 accuracy
// (one half-precision pass would reorder near-ties), so every fp32 value is split
//     x = hi + 2^-11 lo',   hi = RN_fp16(x),   lo' = RN_fp16((x - hi) * 2^11)            (22 significant bits)
//     u . i  ~=  hi_u . hi_i  +  2^-11 (hi_u . lo'_i + lo'_u . hi_i)
// and the two brackets accumulate in two fp32 TMEM accumulators, combined when they are read out.
//
// What changed against eval_umma.cu (3xTF32, 5.4 ms): there, eight producer warps re-gathered, re-split and
// re-stored the USER tile for every item tile, and the grid was one CTA per 128 users (52 CTAs on 148 SMs
// for an 8-way user shard).  Here
//   * eval_split_kernel splits the user rows and the item rows ONCE into fp16 hi / lo' tiles that already
//     have the K-major SWIZZLE_128B shared-memory layout (128 rows x 64 features = 16 KB per array), so a
//     tile is a contiguous blob in HBM / L2;
//   * the main kernel keeps its user tile (hi and lo', all K-blocks, 96 KB at dim 176) RESIDENT in shared
//     memory and streams item K-blocks [I_hi ; I_lo'] (32 KB) through a 3-stage ring with
//     cp.async.bulk + mbarrier complete_tx: no LSU gather, no conversion, no st.shared on the operand path -
//     one thread issues the copies;
//   * per K-step of 16 features: U_hi x [I_hi ; I_lo']^T (N = 256: columns 0..127 hi.hi, 128..255 hi.lo')
//     and U_lo' x I_hi^T (N = 128, into columns 128..255): kind::f16 moves half the operand bytes of TF32 and
//     runs at twice its rate;
//   * work unit = (user tile, item range): the item axis is cut into S ranges so that the grid fills the
//     SMs at any user count (an 8-way user shard of Amazon-book has 53 user tiles); every unit keeps its own
//     top-K and topk_merge_kernel merges the S sorted lists of a user (ties -> lower item index, as a
//     direct insertion would).
// Range: |x| < 65504 (fp16).  eval_split_kernel checks it and raises a device word; the main kernel then does
// nothing and the 3xTF32 kernel (fp32 range) launched behind it computes the result instead - no host round trip.
#include <limits.h>

#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace kgat {
namespace {

constexpr int TM = 128, TN = 128;                 // users per CTA, items per accumulator tile
constexpr int KB = 64;                            // features per K-block: one 128-byte swizzled fp16 row
constexpr int NKB_MAX = 3;                        // dim <= 192: the user tile must stay resident
constexpr uint32_t ARR_BYTES = TM * KB * 2;       // 16 KB: 128 rows x 64 fp16
constexpr uint32_t BLK_BYTES = 2 * ARR_BYTES;     // hi, lo' of one K-block
constexpr int NST = 3;                            // item K-block ring
constexpr int THREADS = 256;                      // warp 0 copies, warp 1 MMA, warps 4..7 top-K
constexpr int N_EPI = 128;
constexpr size_t SMEM_LIMIT = 232448;             // 227 KB opt-in shared memory per CTA
constexpr int CAP = 12, CH = 8;                   // candidate-buffer depth per row, columns per TMEM load
constexpr int S_MAX = 16;                         // item ranges per user tile
constexpr uint32_t TMEM_COLS = 512;               // 2 buffers x (128 + 128) columns
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;

__host__ __device__ inline size_t smem_bytes(int nkb, int K) {
    return static_cast<size_t>(nkb + NST) * BLK_BYTES + static_cast<size_t>(K + CAP) * TM * 8 + 128;
}

// instruction descriptor for kind::f16 with fp16 operands (a_format = b_format = 0), fp32 accumulate
__host__ __device__ constexpr uint32_t make_idesc_f16(int m, int n) {
    return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes)
                 : "memory");
}

// global -> shared bulk copy (bytes % 16 == 0, both addresses 16-byte aligned); completion is signalled on
// the mbarrier as a transaction of `bytes`
__device__ __forceinline__ void bulk_copy_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(umma::smem_u32(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (a, b) -> packed hi pair and packed scaled-residual pair; ok &= both inside the fp16 range
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo, bool& ok) {
    ok = ok && (fabsf(a) < 65504.0f) && (fabsf(b) < 65504.0f);
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = pack_h2(h);
    lo = pack_h2(__floats2half2_rn((a - f.x) * LO_SCALE, (b - f.y) * LO_SCALE));
}

// ------------------------------------------------------------------------------------------------
// operand preparation: one thread = one 16-byte chunk (8 features) of one row of one K-block
// ------------------------------------------------------------------------------------------------
// blob layout: tile t (128 rows), K-block kb: [hi 16 KB][lo' 16 KB] at ((t * nkb + kb) * 32 KB); inside an
// array element (r, p), p = feature index inside the K-block, sits at the SWIZZLE_128B position
// (r / 8) * 1024 + (r % 8) * 128 + (((p / 8) ^ (r % 8)) * 16) + (p % 8) * 2.  Rows past the end and features
// past dim are zero.
__global__ void __launch_bounds__(256)
eval_split_kernel(const float* __restrict__ emb, int64_t stride, int dim,
                  const int32_t* __restrict__ user_ids, int n_users, int u_tiles,
                  const int32_t* __restrict__ item_ids, int n_items, int i_tiles, int nkb,
                  unsigned char* __restrict__ u_blob, unsigned char* __restrict__ i_blob,
                  int* __restrict__ fallback) {
    const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int c = static_cast<int>(gid & 7), r = static_cast<int>((gid >> 3) & 127);
    const int64_t blk = gid >> 10;
    const int kb = static_cast<int>(blk % nkb);
    int tile = static_cast<int>(blk / nkb);
    if (tile >= u_tiles + i_tiles) return;
    const bool is_user = tile < u_tiles;
    if (!is_user) tile -= u_tiles;
    const int row = tile * TM + r;
    const int n_rows = is_user ? n_users : n_items;
    const int32_t* ids = is_user ? user_ids : item_ids;
    const int k0 = kb * KB + c * 8;
    float4 v0 = make_float4(0.f, 0.f, 0.f, 0.f), v1 = v0;
    if (row < n_rows) {
        const float* src = emb + static_cast<size_t>(__ldg(ids + row)) * stride + k0;
        if (k0 < dim) v0 = ldg4(src);
        if (k0 + 4 < dim) v1 = ldg4(src + 4);
    }
    bool ok = true;
    uint4 hi, lo;
    split_pair(v0.x, v0.y, hi.x, lo.x, ok);
    split_pair(v0.z, v0.w, hi.y, lo.y, ok);
    split_pair(v1.x, v1.y, hi.z, lo.z, ok);
    split_pair(v1.z, v1.w, hi.w, lo.w, ok);
    if (!ok) *reinterpret_cast<volatile int*>(fallback) = 1;       // a value outside the fp16 range (or NaN)
    unsigned char* dst = (is_user ? u_blob : i_blob) + (static_cast<size_t>(tile) * nkb + kb) * BLK_BYTES +
                         (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(dst) = hi;
    *reinterpret_cast<uint4*>(dst + ARR_BYTES) = lo;
}

// ------------------------------------------------------------------------------------------------
// scoring + top-K
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 1)
score_topk_f16_kernel(const unsigned char* __restrict__ u_blob, const unsigned char* __restrict__ i_blob,
                      int n_users, int n_items, int nkb, int n_ks, int S,
                      const int32_t* __restrict__ mask_ptr, const int32_t* __restrict__ mask_items, int K,
                      int32_t* __restrict__ out_idx, float* __restrict__ out_score, int* __restrict__ err,
                      const int* __restrict__ fallback) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sU = smem_raw;                                     // nkb x [hi 16 KB][lo' 16 KB]
    unsigned char* sB = sU + static_cast<size_t>(nkb) * BLK_BYTES;    // NST x [I_hi 16 KB][I_lo' 16 KB]
    float* ls = reinterpret_cast<float*>(sB + NST * BLK_BYTES);       // sorted top-K lists, [position][row]
    int* li = reinterpret_cast<int*>(ls + K * TM);
    float* cs = reinterpret_cast<float*>(li + K * TM);                // candidate buffers, [slot][row]
    int* ci = reinterpret_cast<int*>(cs + CAP * TM);
    uint64_t* full = reinterpret_cast<uint64_t*>(ci + CAP * TM);      // [NST]
    uint64_t* empty = full + NST;                                     // [NST]
    uint64_t* acc_full = empty + NST;                                 // [2]
    uint64_t* acc_empty = acc_full + 2;                               // [2]
    uint64_t* u_full = acc_empty + 2;                                 // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_full + 1);
    volatile int* flags = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [0] abort, [1] epilogue abort, [2] skip

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // operands outside the fp16 range: this kernel stands down, the fp32-range kernel behind it runs
    if (tid == 0) flags[2] = *reinterpret_cast<const volatile int*>(fallback);
    __syncthreads();
    if (flags[2]) return;

    if (warp == 1) umma::tmem_alloc(tmem_slot, TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < NST; ++i) {
            umma::mbar_init(&full[i], 1);
            umma::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&acc_full[i], 1);
            umma::mbar_init(&acc_empty[i], N_EPI);
        }
        umma::mbar_init(u_full, 1);
        umma::fence_barrier_init();
        flags[0] = 0;
        flags[1] = 0;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int unit = blockIdx.x, ut = unit / S, sp = unit - ut * S;
    const int n_it = (n_items + TN - 1) / TN;
    const int it0 = static_cast<int>(static_cast<long long>(n_it) * sp / S);
    const int it1 = static_cast<int>(static_cast<long long>(n_it) * (sp + 1) / S);
    const int u0 = ut * TM;

    if (warp == 0) {
        // =========================================================== copy issuer
        if (lane == 0) {
            const unsigned char* usrc = u_blob + static_cast<size_t>(ut) * nkb * BLK_BYTES;
            mbar_expect_tx(u_full, static_cast<uint32_t>(nkb) * BLK_BYTES);
            for (int kb = 0; kb < nkb; ++kb)
                bulk_copy_g2s(sU + static_cast<size_t>(kb) * BLK_BYTES, usrc + static_cast<size_t>(kb) * BLK_BYTES,
                              BLK_BYTES, u_full);
            int slot = 0, use = 0;
            for (int it = it0; it < it1 && !flags[0]; ++it) {
                const unsigned char* isrc = i_blob + static_cast<size_t>(it) * nkb * BLK_BYTES;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (!umma::mbar_wait(&empty[slot], (use & 1) ^ 1)) flags[0] = 1;
                    mbar_expect_tx(&full[slot], BLK_BYTES);
                    bulk_copy_g2s(sB + static_cast<size_t>(slot) * BLK_BYTES, isrc + static_cast<size_t>(kb) * BLK_BYTES,
                                  BLK_BYTES, &full[slot]);
                    if (++slot == NST) { slot = 0; ++use; }
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc256 = make_idesc_f16(TM, 256), idesc128 = make_idesc_f16(TM, 128);
            if (!umma::mbar_wait(u_full, 0)) flags[0] = 1;
            umma::fence_after_sync();
            const uint32_t aU0 = umma::smem_u32(sU), bI0 = umma::smem_u32(sB);
            int slot = 0, use = 0;
            for (int it = it0; it < it1 && !flags[0]; ++it) {
                const int n = it - it0, buf = n & 1;
                if (!umma::mbar_wait(&acc_empty[buf], ((n >> 1) & 1) ^ 1)) flags[0] = 1;
                umma::fence_after_sync();
                const uint32_t tD = tmem + buf * 256;
                for (int kb = 0; kb < nkb; ++kb) {
                    if (!umma::mbar_wait(&full[slot], use & 1)) flags[0] = 1;
                    umma::fence_after_sync();
                    const uint32_t aU = aU0 + kb * BLK_BYTES, bI = bI0 + slot * BLK_BYTES;
                    const int nks = min(KB / 16, n_ks - kb * (KB / 16));     // the last K-block may be partial
                    for (int ks = 0; ks < nks; ++ks) {
                        const uint32_t off = ks * 32;                          // 16 fp16 inside the 128-byte row
                        const uint64_t dB = umma::make_smem_desc_sw128(bI + off);   // rows 0..127 I_hi, 128..255 I_lo'
                        mma_f16(tD, umma::make_smem_desc_sw128(aU + off), dB, idesc256, (kb | ks) != 0);
                        mma_f16(tD + 128, umma::make_smem_desc_sw128(aU + ARR_BYTES + off), dB, idesc128, 1);
                    }
                    umma::mma_commit(&empty[slot]);
                    if (++slot == NST) { slot = 0; ++use; }
                }
                umma::mma_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 4) {
        // =========================================================== epilogue (warps 4..7): mask + top-K
        const int q = warp & 3, row = q * 32 + lane, et = tid - 128;
        const bool sel = u0 + row < n_users;
        const int item_begin = it0 * TN;
        int cnt = 0, mp = 0, mend = 0, next_mask = INT_MAX;
        float thr = -INFINITY;
        if (sel && mask_ptr && mask_items) {
            mp = mask_ptr[u0 + row];
            mend = mask_ptr[u0 + row + 1];
            while (mp < mend && mask_items[mp] < item_begin) ++mp;      // this unit starts inside the item axis
            next_mask = mp < mend ? mask_items[mp] : INT_MAX;
        }
        float* const lsr = ls + row;                     // element p of this row's list: lsr[p * TM]
        int* const lir = li + row;
        float* const csr = cs + row;
        int* const cir = ci + row;
        int ncand = 0;
        // merge the buffered candidates, in arrival (= ascending item) order, into the sorted list; a
        // candidate displaces only entries it beats strictly, so ties keep the lower index first
        auto flush = [&]() {
            for (int t = 0; t < ncand; ++t) {
                const float sc = csr[t * TM];
                if (cnt < K || sc > thr) {
                    int p = cnt < K ? cnt : K - 1;
                    while (p > 0 && lsr[(p - 1) * TM] < sc) {
                        lsr[p * TM] = lsr[(p - 1) * TM];
                        lir[p * TM] = lir[(p - 1) * TM];
                        --p;
                    }
                    lsr[p * TM] = sc;
                    lir[p * TM] = cir[t * TM];
                    if (cnt < K) ++cnt;
                    if (cnt == K) thr = lsr[(K - 1) * TM];
                }
            }
            ncand = 0;
        };
        auto consume = [&](const float* a, const float* b, int item0, bool tile_full) {
            // when ANY row of the warp is about to overflow its buffer every row merges now (rows of a warp
            // fill their buffers at similar rates, so the serial insertion chain runs a few times per tile)
            if (__any_sync(FULL_MASK, sel && ncand > CAP - CH)) flush();
            if (!sel) return;
            if (tile_full && cnt == K && next_mask >= item0 + CH) {
                // fast path (almost every chunk): no bounds, no mask, list already full
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const float sc = fmaf(b[i], LO_INV, a[i]);
                    if (sc > thr) {
                        csr[ncand * TM] = sc;
                        cir[ncand * TM] = item0 + i;
                        ++ncand;
                    }
                }
            } else {
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    const int item = item0 + i;
                    float sc = fmaf(b[i], LO_INV, a[i]);
                    if (item == next_mask) {                // train item: score forced to 0.0 (metric.py:50)
                        sc = 0.0f;
                        do {
                            ++mp;
                            next_mask = mp < mend ? mask_items[mp] : INT_MAX;
                        } while (next_mask == item);
                    }
                    if (item < n_items && (cnt < K || sc > thr)) {
                        csr[ncand * TM] = sc;
                        cir[ncand * TM] = item;
                        ++ncand;
                    }
                }
            }
        };
        for (int it = it0; it < it1; ++it) {
            const int n = it - it0, buf = n & 1;
            if (!umma::mbar_wait(&acc_full[buf], (n >> 1) & 1)) { flags[0] = 1; flags[1] = 1; }
            if (et == 0 && flags[0]) flags[1] = 1;
            umma::fence_after_sync();
            asm volatile("bar.sync 1, %0;" ::"n"(N_EPI) : "memory");
            if (flags[1]) break;
            const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
            const bool tile_full = (it + 1) * TN <= n_items;
            const int item_tile = it * TN;
            float a0[CH], b0[CH], a1[CH], b1[CH];            // chunk c + 1 loads while chunk c is consumed
            umma::tmem_ld8(ta, a0);
            umma::tmem_ld8(ta + 128, b0);
#pragma unroll 1
            for (int c0 = 0; c0 < TN; c0 += 2 * CH) {
                umma::tmem_ld_wait();
                umma::tmem_ld8(ta + c0 + CH, a1);
                umma::tmem_ld8(ta + 128 + c0 + CH, b1);
                consume(a0, b0, item_tile + c0, tile_full);
                umma::tmem_ld_wait();
                if (c0 + 2 * CH < TN) {
                    umma::tmem_ld8(ta + c0 + 2 * CH, a0);
                    umma::tmem_ld8(ta + 128 + c0 + 2 * CH, b0);
                }
                consume(a1, b1, item_tile + c0 + CH, tile_full);
            }
            umma::fence_before_sync();
            umma::mbar_arrive(&acc_empty[buf]);
        }
        if (sel && !flags[1]) {
            flush();
            const size_t o = (static_cast<size_t>(u0 + row) * S + sp) * K;
            for (int p = 0; p < K; ++p) {
                const bool ok = p < cnt;
                out_idx[o + p] = ok ? lir[p * TM] : -1;
                if (out_score) out_score[o + p] = ok ? lsr[p * TM] : -INFINITY;
            }
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0 && flags[0]) atomicExch(err, 1);
    if (warp == 1) umma::tmem_dealloc(tmem, TMEM_COLS);
}

// S sorted lists per user (one per item range, ascending ranges) -> the user's top-K.  Equal scores keep the
// entry of the lower range first and inside a range the lists are already index-ordered, so ties rank the
// lower item index first, exactly as one list over the whole item axis would.
__global__ void __launch_bounds__(128)
topk_merge_kernel(const int32_t* __restrict__ part_idx, const float* __restrict__ part_score, int n_users, int S,
                  int K, int32_t* __restrict__ out_idx, float* __restrict__ out_score,
                  const int* __restrict__ fallback) {
    if (*reinterpret_cast<const volatile int*>(fallback)) return;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    unsigned char ptr[S_MAX];
#pragma unroll
    for (int s = 0; s < S_MAX; ++s) ptr[s] = 0;
    const size_t base = static_cast<size_t>(u) * S * K;
    for (int k = 0; k < K; ++k) {
        int best = -1, best_idx = -1;
        float best_sc = -INFINITY;
        for (int s = 0; s < S; ++s) {
            if (ptr[s] >= K) continue;
            const size_t o = base + static_cast<size_t>(s) * K + ptr[s];
            const int idx = part_idx[o];
            if (idx < 0) continue;
            const float sc = part_score[o];
            if (best < 0 || sc > best_sc) { best = s; best_sc = sc; best_idx = idx; }
        }
        out_idx[static_cast<size_t>(u) * K + k] = best_idx;
        if (out_score) out_score[static_cast<size_t>(u) * K + k] = best >= 0 ? best_sc : -INFINITY;
        if (best >= 0) ++ptr[best];
    }
}

__device__ int g_eval_f16_error = 0;      // a tensor-core / copy completion barrier timed out
__device__ int g_eval_fallback[2] = {0, 0};   // [0] operands outside the fp16 range (per call), [1] fallbacks taken

inline size_t align1k(size_t x) { return (x + 1023) / 1024 * 1024; }

// item ranges per user tile: fill the SMs at any user count without paying the user-tile load and the
// pipeline ramp (~3 item tiles) more often than it helps
inline int pick_splits(int u_tiles, int n_it, int sms) {
    int best_s = 1;
    long long best_cost = -1;
    for (int s = 1; s <= S_MAX && s <= n_it; ++s) {
        const long long waves = (static_cast<long long>(u_tiles) * s + sms - 1) / sms;
        const long long cost = waves * ((n_it + s - 1) / s + 3);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s; }
    }
    return best_s;
}

struct Plan { int u_tiles, i_tiles, nkb, n_ks, S; size_t u_blob, i_blob, part_idx, part_score, total; };

inline Plan make_plan(int64_t n_users, int64_t n_items, int dim, int K, int sms) {
    Plan p{};
    p.u_tiles = static_cast<int>((n_users + TM - 1) / TM);
    p.i_tiles = static_cast<int>((n_items + TN - 1) / TN);
    p.nkb = (dim + KB - 1) / KB;
    p.n_ks = (dim + 15) / 16;
    p.S = pick_splits(p.u_tiles, p.i_tiles, sms);
    size_t off = 0;
    p.u_blob = off; off += align1k(static_cast<size_t>(p.u_tiles) * p.nkb * BLK_BYTES);
    p.i_blob = off; off += align1k(static_cast<size_t>(p.i_tiles) * p.nkb * BLK_BYTES);
    p.part_idx = off; off += align1k(static_cast<size_t>(n_users) * S_MAX * K * 4);
    p.part_score = off; off += align1k(static_cast<size_t>(n_users) * S_MAX * K * 4);
    p.total = off;
    return p;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

// dim <= 192 (resident user tile) and a top-K whose lists fit next to it: K <= 22 at dim 176, K <= 54 at dim <= 128
bool kgat_score_topk_f16_supported(int dim, int K) {
    return dim > 0 && dim % 4 == 0 && dim <= NKB_MAX * KB && K > 0 && K <= 64 &&
           smem_bytes((dim + KB - 1) / KB, K) <= SMEM_LIMIT;
}

size_t kgat_score_topk_f16_workspace(int64_t n_users, int64_t n_items, int dim, int K) {
    return make_plan(n_users, n_items, dim, K, 148).total;       // the partial-list area does not depend on S
}

int kgat_eval_f16_error_take() {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_eval_f16_error, sizeof(int)) != cudaSuccess) return KGAT_ERR_CUDA;
    if (v != 0) cudaMemcpyToSymbol(g_eval_f16_error, &zero, sizeof(int));
    return v;
}

int kgat_eval_fallbacks_take() {
    int v[2] = {0, 0}, zero = 0;
    if (cudaMemcpyFromSymbol(v, g_eval_fallback, sizeof(v)) != cudaSuccess) return KGAT_ERR_CUDA;
    if (v[1] != 0 && cudaMemcpyToSymbol(g_eval_fallback, &zero, sizeof(int), sizeof(int)) != cudaSuccess)
        return KGAT_ERR_CUDA;
    return v[1];
}

int kgat_launch_score_topk_umma(const float* emb, int64_t emb_stride, int dim, const int32_t* user_ids, int n_users,
                                const int32_t* item_ids, int n_items, const int32_t* mask_ptr,
                                const int32_t* mask_items, int K, int32_t* topk_idx, float* topk_score,
                                int* only_if, cudaStream_t st);

// split -> score + top-K per (user tile, item range) -> merge; the 3xTF32 kernel runs behind iff an operand
// was outside the fp16 range (an empty launch otherwise)
int kgat_launch_score_topk_f16(const float* emb, int64_t emb_stride, int dim, const int32_t* user_ids, int n_users,
                               const int32_t* item_ids, int n_items, const int32_t* mask_ptr,
                               const int32_t* mask_items, int K, int32_t* topk_idx, float* topk_score,
                               void* workspace, size_t workspace_bytes, cudaStream_t st) {
    static PerDevice<int*> err_words, fb_words;
    static PerDevice<size_t> smem_set;
    int*& err_word = err_words.get();
    int*& fb_word = fb_words.get();
    if (!err_word) KGAT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&err_word), g_eval_f16_error));
    if (!fb_word) KGAT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&fb_word), g_eval_fallback));
    const Plan p = make_plan(n_users, n_items, dim, K, num_sms());
    if (workspace_bytes < p.total) return KGAT_ERR_WORKSPACE;
    const size_t smem = smem_bytes(p.nkb, K);
    if (smem > smem_set.get()) {
        KGAT_CUDA(cudaFuncSetAttribute(score_topk_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set.get() = smem;
    }
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    KGAT_CUDA(cudaMemsetAsync(fb_word, 0, sizeof(int), st));
    const int64_t chunks = static_cast<int64_t>(p.u_tiles + p.i_tiles) * p.nkb * 1024;
    eval_split_kernel<<<static_cast<unsigned>((chunks + 255) / 256), 256, 0, st>>>(
        emb, emb_stride, dim, user_ids, n_users, p.u_tiles, item_ids, n_items, p.i_tiles, p.nkb, ws + p.u_blob,
        ws + p.i_blob, fb_word);
    KGAT_LAUNCH_CHECK();
    int32_t* pidx = p.S > 1 ? reinterpret_cast<int32_t*>(ws + p.part_idx) : topk_idx;
    float* pscore = p.S > 1 ? reinterpret_cast<float*>(ws + p.part_score) : topk_score;
    score_topk_f16_kernel<<<static_cast<unsigned>(p.u_tiles) * p.S, THREADS, smem, st>>>(
        ws + p.u_blob, ws + p.i_blob, n_users, n_items, p.nkb, p.n_ks, p.S, mask_ptr, mask_items, K, pidx, pscore,
        err_word, fb_word);
    KGAT_LAUNCH_CHECK();
    if (p.S > 1) {
        topk_merge_kernel<<<(n_users + 127) / 128, 128, 0, st>>>(pidx, pscore, n_users, p.S, K, topk_idx, topk_score,
                                                                 fb_word);
        KGAT_LAUNCH_CHECK();
    }
    return kgat_launch_score_topk_umma(emb, emb_stride, dim, user_ids, n_users, item_ids, n_items, mask_ptr, mask_items, K,
                                       topk_idx, topk_score, fb_word, st);
}
