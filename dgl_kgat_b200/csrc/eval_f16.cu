// (5) Full-rank user x item scoring on the tensor cores, split-fp16 operands prepared once per call and
// streamed by bulk asynchronous copies, with the per-row top-K fused into the accumulator read-out.
//
// Replaces the per-user loop of the reference (metric.py:44-59): item-matrix re-gather, GEMV, mask, full
// sort, device->host sync per user.  score[u, i] = <emb[user u], emb[item i]> is a (U x dim) x (dim x I)
// GEMM (469 GFLOP for the 53 469 eval users of the Amazon-book shape); the ranking needs fp32 accuracy
// (one half-precision pass would reorder near-ties), so every fp32 value is pre-scaled by one power of two
// s (chosen so that max |x| s lies in [2^13, 2^14): the residual below stays a normal fp16 number) and split
//     x s = hi + lo,   hi = RN_fp16(x s),   lo = RN_fp16(x s - hi)                       (22 significant bits)
//     s^2 u . i  ~=  hi_u . hi_i + hi_u . lo_i + lo_u . hi_i
// all three products into ONE fp32 TMEM accumulator; the ranking does not see the scale.
//
// What changed against eval_umma.cu (3xTF32, 5.4 ms): there, eight producer warps re-gathered, re-split and
// re-stored the USER tile for every item tile, and the grid was one CTA per 128 users (52 CTAs on 148 SMs
// for an 8-way user shard).  Here
//   * eval_split_kernel splits the user rows and the item rows ONCE into fp16 hi / lo' tiles that already
//     have the K-major SWIZZLE_128B shared-memory layout (128 rows x 64 features = 16 KB per array), so a
//     tile is a contiguous blob in HBM / L2;
//   * the main kernel keeps its user tile (hi and lo, all K-blocks, 96 KB at dim 176) RESIDENT in shared
//     memory and streams the item arrays (I_hi, I_lo: 16 KB each) through a 4-slot ring with
//     cp.async.bulk + mbarrier complete_tx: no LSU gather, no conversion, no st.shared on the operand path -
//     one thread issues the copies;
//   * per K-step of 16 features three M = 128, N = 128 kind::f16 MMAs (half the operand bytes of TF32 at
//     twice its rate) into one of FOUR accumulator buffers;
//   * two read-out groups of four warps alternate over the item tiles (one warp per SM sub-partition could
//     not keep up with the tensor pipe: measured 13 000 clocks per tile against 2 300), each with its own
//     candidate buffers and sorted lists;
//   * work unit = (user tile, item range): the item axis is cut into S ranges so that the grid fills the
//     SMs at any user count (an 8-way user shard of Amazon-book has 53 user tiles); every unit keeps its own
//     top-K and topk_merge_kernel merges the sorted lists of a user (ties -> lower item index, as a
//     direct insertion would).
// Range: any finite fp32 magnitude (the scale is computed per call).  inf / NaN or a dynamic range above 2^16
// inside the operands raise a device word; the main kernel then does nothing and the 3xTF32 kernel launched
// behind it computes the result instead - no host round trip.
#include <limits.h>

#include <cstdlib>

#include <cuda_fp16.h>

#include "common.cuh"
#include "umma.cuh"

namespace kgat {
namespace {

constexpr int TM = 128, TN = 128;                 // users per CTA, items per accumulator tile
constexpr int KB = 64;                            // features per K-block: one 128-byte swizzled fp16 row
constexpr int NKB_MAX = 3;                        // dim <= 192: the user tile must stay resident
constexpr uint32_t ARR_BYTES = TM * KB * 2;       // 16 KB: 128 rows x 64 fp16
constexpr uint32_t BLK_BYTES = 2 * ARR_BYTES;     // hi, lo of one K-block
constexpr int NSLOT = 4;                          // ring of item ARRAYS (I_hi and I_lo are consumed separately)
constexpr int N_GRP = 2;                          // read-out groups: group g takes the tiles n with n % 2 == g
constexpr int N_EPI = 128;                        // threads per read-out group (one TMEM lane each)
constexpr int THREADS = 128 + N_GRP * N_EPI;      // warp 0 copies, warp 1 MMA, warps 4..7 / 8..11 top-K
constexpr int N_ACC = 4;                          // accumulator buffers (128 TMEM columns each)
constexpr size_t SMEM_LIMIT = 232448;             // 227 KB opt-in shared memory per CTA
constexpr int CAP = 12, CH = 8;                   // candidate-buffer depth per row, columns per chunk
constexpr int S_MAX = 8;                          // item ranges per user tile
constexpr int L_MAX = N_GRP * S_MAX;              // sorted lists per user that the merge sees
constexpr uint32_t TMEM_COLS = 512;
constexpr int TARGET_EXP = 13;                    // operands are pre-scaled so that max |x| lies in [2^13, 2^14)

__host__ __device__ inline size_t smem_bytes(int nkb, int K) {
    return static_cast<size_t>(nkb) * BLK_BYTES + static_cast<size_t>(NSLOT) * ARR_BYTES +
           static_cast<size_t>(N_GRP) * (K + CAP) * TM * 8 + N_GRP * TM * 4 + 256;
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

// 64-bit ranking key: larger key = ranks earlier.  High word: the score's bits made order-preserving as an
// unsigned integer; low word: ~index, so that among equal scores the LOWER index has the larger key.
__device__ __forceinline__ unsigned long long make_key(float score, int idx) {
    const uint32_t b = __float_as_uint(score);
    const uint32_t o = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return (static_cast<unsigned long long>(o) << 32) | static_cast<uint32_t>(~idx);
}
__device__ __forceinline__ float key_score(unsigned long long k) {
    const uint32_t o = static_cast<uint32_t>(k >> 32);
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
__device__ __forceinline__ int key_index(unsigned long long k) { return static_cast<int>(~static_cast<uint32_t>(k)); }

__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (a, b) already scaled -> packed hi pair and packed residual pair
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = pack_h2(h);
    lo = pack_h2(__floats2half2_rn(a - f.x, b - f.y));
}

// ------------------------------------------------------------------------------------------------
// operand preparation: one thread = one 16-byte chunk (8 features) of one row of one K-block
// ------------------------------------------------------------------------------------------------
struct PrepIdx { int c, r, kb, tile; bool is_user, valid; };

__device__ __forceinline__ PrepIdx prep_index(int u_tiles, int i_tiles, int nkb) {
    const int64_t gid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    PrepIdx p;
    p.c = static_cast<int>(gid & 7);
    p.r = static_cast<int>((gid >> 3) & 127);
    const int64_t blk = gid >> 10;
    p.kb = static_cast<int>(blk % nkb);
    p.tile = static_cast<int>(blk / nkb);
    p.valid = p.tile < u_tiles + i_tiles;
    p.is_user = p.tile < u_tiles;
    if (!p.is_user) p.tile -= u_tiles;
    return p;
}

__device__ __forceinline__ void prep_load(const PrepIdx& p, const float* __restrict__ emb, int64_t stride, int dim,
                                          const int32_t* __restrict__ user_ids, int n_users,
                                          const int32_t* __restrict__ item_ids, int n_items, float4& v0, float4& v1) {
    v0 = make_float4(0.f, 0.f, 0.f, 0.f);
    v1 = v0;
    const int row = p.tile * TM + p.r;
    const int n_rows = p.is_user ? n_users : n_items;
    const int32_t* ids = p.is_user ? user_ids : item_ids;
    const int k0 = p.kb * KB + p.c * 8;
    if (p.valid && row < n_rows) {
        const float* src = emb + static_cast<size_t>(__ldg(ids + row)) * stride + k0;
        if (k0 < dim) v0 = ldg4(src);
        if (k0 + 4 < dim) v1 = ldg4(src + 4);
    }
}

// pass 1: word[0] = bits of max |x| over the user and item rows (non-negative floats order like their bit
// patterns), word[1] |= 1 if a value is inf / NaN (then the fp32 kernel takes over), word[3] = bits of the
// smallest non-zero chunk maximum (8 consecutive features of a row): the dynamic range of the operands
__global__ void __launch_bounds__(256)
eval_absmax_kernel(const float* __restrict__ emb, int64_t stride, int dim,
                   const int32_t* __restrict__ user_ids, int n_users, int u_tiles,
                   const int32_t* __restrict__ item_ids, int n_items, int i_tiles, int nkb,
                   unsigned int* __restrict__ word) {
    const PrepIdx p = prep_index(u_tiles, i_tiles, nkb);
    float4 v0, v1;
    prep_load(p, emb, stride, dim, user_ids, n_users, item_ids, n_items, v0, v1);
    float m = fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                    fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w))));
    const float sum = v0.x + v0.y + v0.z + v0.w + v1.x + v1.y + v1.z + v1.w;       // NaN / inf propagate
    bool bad = !(fabsf(sum) < INFINITY) || !(m < INFINITY);
    m = warp_max(m);
    bad = __any_sync(FULL_MASK, bad);
    float lo = m > 0.f ? m : INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lo = fminf(lo, __shfl_xor_sync(FULL_MASK, lo, o));
    // one set of atomics per CTA, and only when it would change the word (the words move monotonically, so a stale
    // read can only cause a redundant atomic): a per-warp atomic on one address serialised 58 000 of them - 85 us
    // for a 56 MB read
    __shared__ float s_m[8], s_lo[8];
    __shared__ int s_bad[8];
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_m[warp] = m; s_lo[warp] = lo; s_bad[warp] = bad ? 1 : 0; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int any_bad = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) { m = fmaxf(m, s_m[w]); lo = fminf(lo, s_lo[w]); any_bad |= s_bad[w]; }
        const volatile unsigned int* vw = word;
        if (m > 0.f && __float_as_uint(m) > vw[0]) atomicMax(word, __float_as_uint(m));
        if (lo < INFINITY && __float_as_uint(lo) < vw[3]) atomicMin(word + 3, __float_as_uint(lo));
        if (any_bad) atomicOr(word + 1, 1u);
    }
}

// The fp16 pair resolves a value x to max(2^-22 |x|, 2^-39 max|x|): fine while the embeddings span less than
// ~2^16 in magnitude.  A table with a wilder dynamic range (an outlier 10^5 times the typical entry) goes to
// the fp32-range kernel instead.
constexpr float MAX_DYNAMIC_RANGE = 65536.0f;

// scale = 2^e with max |x| * scale in [2^13, 2^14): the fp16 pair (hi, lo) = (RN(x s), RN(x s - hi)) then
// carries 22 significant bits for every value within 2^-16 of the maximum and <= 2^-38 max ABSOLUTE error
// below that, whatever the magnitude of the embeddings - no range limit, no separate scale for the residual
__device__ __forceinline__ float operand_scale(unsigned int max_bits) {
    const float m = __uint_as_float(max_bits);
    if (!(m > 0.f) || !(m < INFINITY)) return 1.0f;
    return ldexpf(1.0f, TARGET_EXP - ilogbf(m));
}

// pass 2.  blob layout: tile t (128 rows), K-block kb: [hi 16 KB][lo 16 KB] at ((t * nkb + kb) * 32 KB);
// inside an array element (r, p), p = feature index inside the K-block, sits at the SWIZZLE_128B position
// (r / 8) * 1024 + (r % 8) * 128 + (((p / 8) ^ (r % 8)) * 16) + (p % 8) * 2.  Rows past the end and features
// past dim are zero.
__global__ void __launch_bounds__(256)
eval_split_kernel(const float* __restrict__ emb, int64_t stride, int dim,
                  const int32_t* __restrict__ user_ids, int n_users, int u_tiles,
                  const int32_t* __restrict__ item_ids, int n_items, int i_tiles, int nkb,
                  unsigned char* __restrict__ u_blob, unsigned char* __restrict__ i_blob,
                  const unsigned int* word) {
    const PrepIdx p = prep_index(u_tiles, i_tiles, nkb);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        // the decision every later kernel of this call reads (stream order): word[1] != 0 -> fp32 kernel
        const float hi = __uint_as_float(word[0]), lo = __uint_as_float(word[3]);
        if (hi > 0.f && lo < INFINITY && hi > lo * MAX_DYNAMIC_RANGE) atomicOr(const_cast<unsigned int*>(word) + 1, 1u);
    }
    if (!p.valid) return;
    float4 v0, v1;
    prep_load(p, emb, stride, dim, user_ids, n_users, item_ids, n_items, v0, v1);
    const float sc = operand_scale(word[0]);
    uint4 hi, lo;
    split_pair(v0.x * sc, v0.y * sc, hi.x, lo.x);
    split_pair(v0.z * sc, v0.w * sc, hi.y, lo.y);
    split_pair(v1.x * sc, v1.y * sc, hi.z, lo.z);
    split_pair(v1.z * sc, v1.w * sc, hi.w, lo.w);
    unsigned char* dst = (p.is_user ? u_blob : i_blob) + (static_cast<size_t>(p.tile) * nkb + p.kb) * BLK_BYTES +
                         (p.r >> 3) * 1024 + (p.r & 7) * 128 + ((p.c ^ (p.r & 7)) << 4);
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
                      const unsigned int* __restrict__ word) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sU = smem_raw;                                     // nkb x [hi 16 KB][lo 16 KB]
    unsigned char* sB = sU + static_cast<size_t>(nkb) * BLK_BYTES;    // NSLOT x 16 KB
    unsigned long long* lk_all = reinterpret_cast<unsigned long long*>(sB + NSLOT * ARR_BYTES);   // per group: top-K keys [K][TM]
    float* cs_all = reinterpret_cast<float*>(lk_all + N_GRP * K * TM);    // per group: candidates [CAP][TM]
    int* ci_all = reinterpret_cast<int*>(cs_all + N_GRP * CAP * TM);
    float* thr_pub = reinterpret_cast<float*>(ci_all + N_GRP * CAP * TM);     // [N_GRP][TM] K-th best score per group and row
    uint64_t* full = reinterpret_cast<uint64_t*>(thr_pub + N_GRP * TM);       // [NSLOT]
    uint64_t* empty = full + NSLOT;                                   // [NSLOT]
    uint64_t* acc_full = empty + NSLOT;                               // [N_ACC]
    uint64_t* acc_empty = acc_full + N_ACC;                           // [N_ACC]
    uint64_t* u_full = acc_empty + N_ACC;                             // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(u_full + 1);
    volatile int* flags = reinterpret_cast<volatile int*>(tmem_slot + 1);   // [0] abort, [1] skip

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // inf / NaN among the operands: this kernel stands down, the fp32 kernel behind it runs
    if (tid == 0) flags[1] = static_cast<int>(*reinterpret_cast<const volatile unsigned int*>(word + 1));
    __syncthreads();
    if (flags[1]) return;

    if (warp == 1) umma::tmem_alloc(tmem_slot, TMEM_COLS);
    if (tid == 0) {
        for (int i = 0; i < NSLOT; ++i) {
            umma::mbar_init(&full[i], 1);
            umma::mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < N_ACC; ++i) {
            umma::mbar_init(&acc_full[i], 1);
            umma::mbar_init(&acc_empty[i], N_EPI);
        }
        umma::mbar_init(u_full, 1);
        umma::fence_barrier_init();
        flags[0] = 0;
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
            // the item blob of tiles it0..it1 is one contiguous sequence of arrays: hi, lo, hi, lo, ...
            const unsigned char* isrc = i_blob + static_cast<size_t>(it0) * nkb * BLK_BYTES;
            const int n_arr = (it1 - it0) * nkb * 2;
            int slot = 0, use = 0;
            for (int a = 0; a < n_arr && !flags[0]; ++a) {
                if (!umma::mbar_wait(&empty[slot], (use & 1) ^ 1)) flags[0] = 1;
                mbar_expect_tx(&full[slot], ARR_BYTES);
                bulk_copy_g2s(sB + static_cast<size_t>(slot) * ARR_BYTES, isrc + static_cast<size_t>(a) * ARR_BYTES,
                              ARR_BYTES, &full[slot]);
                if (++slot == NSLOT) { slot = 0; ++use; }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer
        // per K-step of 16 features, all three products into ONE accumulator (same shape, same columns):
        //   I_hi array:  D += U_hi I_hi^T,  D += U_lo I_hi^T        I_lo array:  D += U_hi I_lo^T
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(TM, TN);
            if (!umma::mbar_wait(u_full, 0)) flags[0] = 1;
            umma::fence_after_sync();
            const uint32_t aU0 = umma::smem_u32(sU), bI0 = umma::smem_u32(sB);
            int slot = 0, use = 0;
            for (int it = it0; it < it1 && !flags[0]; ++it) {
                const int n = it - it0, buf = n % N_ACC;
                if (!umma::mbar_wait(&acc_empty[buf], ((n / N_ACC) & 1) ^ 1)) flags[0] = 1;
                umma::fence_after_sync();
                const uint32_t tD = tmem + buf * TN;
                for (int kb = 0; kb < nkb; ++kb) {
                    const uint32_t aHi = aU0 + kb * BLK_BYTES, aLo = aHi + ARR_BYTES;
                    const int nks = min(KB / 16, n_ks - kb * (KB / 16));     // the last K-block may be partial
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if (!umma::mbar_wait(&full[slot], use & 1)) flags[0] = 1;
                        umma::fence_after_sync();
                        const uint32_t bI = bI0 + slot * ARR_BYTES;
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint32_t off = ks * 32;                      // 16 fp16 inside the 128-byte row
                            const uint64_t dB = umma::make_smem_desc_sw128(bI + off);
                            mma_f16(tD, umma::make_smem_desc_sw128(aHi + off), dB, idesc, (kb | ks | h) != 0);
                            if (h == 0) mma_f16(tD, umma::make_smem_desc_sw128(aLo + off), dB, idesc, 1);
                        }
                        umma::mma_commit(&empty[slot]);
                        if (++slot == NSLOT) { slot = 0; ++use; }
                    }
                }
                umma::mma_commit(&acc_full[buf]);
            }
        }
    } else if (warp >= 4) {
        // =========================================================== read-out groups: mask + top-K
        const int grp = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane;
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
        unsigned long long* const lkr = lk_all + grp * K * TM + row;   // entry p of this row's list: lkr[p * TM]
        float* const csr = cs_all + grp * CAP * TM + row;
        int* const cir = ci_all + grp * CAP * TM + row;
        int ncand = 0, minpos = 0;
        // Both groups rank the SAME rows over disjoint tiles; the final list is the best K of their union, so a
        // score below the OTHER group's K-th best can be dropped as well.  thr_own: K-th best of this group's list
        // (candidates must beat it strictly: they come later in item order); thr: what the fast path tests,
        // max(thr_own, just below the other group's published K-th best) - an equal score may still win its
        // tie by index, so it is kept.
        float thr_own = -INFINITY;
        volatile float* const pub_mine = thr_pub + grp * TM + row;
        volatile float* const pub_other = thr_pub + (grp ^ 1) * TM + row;
        *pub_mine = -INFINITY;
        asm volatile("bar.sync 1, %0;" ::"n"(N_GRP * N_EPI) : "memory");     // both groups have initialised
        // The row's K best so far are an UNSORTED set of 64-bit keys (order-preserving score bits, then the
        // complemented item index: the smallest key is the entry that ranks last - lowest score, and among
        // equal scores the highest index).  A candidate that beats the threshold replaces that entry and the
        // minimum is found again by one pass over the K keys: independent loads, no serial shifting (the
        // sorted insertion it replaces was 30 % of the kernel).  Candidates arrive in ascending item order and
        // displace only entries they beat strictly, so ties keep the lower index; the list is sorted once, at the end.
        auto rescan = [&]() {
            unsigned long long mk = lkr[0];
            int mp_ = 0;
#pragma unroll 4
            for (int p = 1; p < K; ++p) {
                const unsigned long long k = lkr[p * TM];
                if (k < mk) { mk = k; mp_ = p; }
            }
            minpos = mp_;
            thr_own = key_score(mk);
            *pub_mine = thr_own;
        };
        auto refresh_thr = [&]() {
            const float o = *pub_other;
            thr = fmaxf(thr_own, o > -INFINITY ? nextafterf(o, -INFINITY) : o);
        };
        auto flush = [&]() {
            for (int t = 0; t < ncand; ++t) {
                const float sc = csr[t * TM];
                if (cnt < K) {
                    lkr[cnt * TM] = make_key(sc, cir[t * TM]);
                    if (++cnt == K) rescan();
                } else if (sc > thr_own) {
                    lkr[minpos * TM] = make_key(sc, cir[t * TM]);
                    rescan();
                }
            }
            ncand = 0;
            refresh_thr();
        };
        // One warp reads out 32 rows, so the steady state has to be cheap in ISSUE SLOTS: a 3-level max tree
        // and one compare decide whether any of the chunk's 8 scores can enter this row's list; only then are
        // the scores looked at one by one.  (Testing and appending every score individually cost ~20
        // instructions per score: 13 000 clocks per tile against the 2 300 the tensor pipe needs - measured.)
        auto consume = [&](const float* sc, int item0, bool tile_full) {
            if (sel) {
                if (tile_full && cnt == K && next_mask >= item0 + CH) {
                    // fast path (almost every chunk): no bounds, no mask, list already full
                    const float mx = fmaxf(fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3])),
                                           fmaxf(fmaxf(sc[4], sc[5]), fmaxf(sc[6], sc[7])));
                    if (mx > thr) {
#pragma unroll
                        for (int i = 0; i < CH; ++i) {
                            if (sc[i] > thr) {
                                csr[ncand * TM] = sc[i];
                                cir[ncand * TM] = item0 + i;
                                ++ncand;
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < CH; ++i) {
                        const int item = item0 + i;
                        float v = sc[i];
                        if (item == next_mask) {            // train item: score forced to 0.0 (metric.py:50)
                            v = 0.0f;
                            do {
                                ++mp;
                                next_mask = mp < mend ? mask_items[mp] : INT_MAX;
                            } while (next_mask == item);
                        }
                        if (item < n_items && (cnt < K || v > thr)) {
                            csr[ncand * TM] = v;
                            cir[ncand * TM] = item;
                            ++ncand;
                        }
                    }
                }
            }
            // when ANY row of the warp could overflow its buffer with the next chunk every row merges now
            if (__any_sync(FULL_MASK, sel && ncand > CAP - CH)) flush();
        };
        for (int it = it0 + grp; it < it1; it += N_GRP) {
            const int n = it - it0, buf = n % N_ACC;
            // no barrier between the read-out warps: each follows the accumulator barriers on its own
            if (!umma::mbar_wait(&acc_full[buf], (n / N_ACC) & 1)) flags[0] = 1;
            umma::fence_after_sync();
            if (flags[0]) break;
            const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * TN;
            const bool tile_full = (it + 1) * TN <= n_items;
            const int item_tile = it * TN;
            if (next_mask < item_tile) {                      // train items inside the other group's tiles
                while (mp < mend && mask_items[mp] < item_tile) ++mp;
                next_mask = mp < mend ? mask_items[mp] : INT_MAX;
            }
            float a0[16], a1[16];                            // columns c + 16.. load while columns c.. are consumed
            umma::tmem_ld16(ta, a0);
#pragma unroll 1
            for (int c0 = 0; c0 < TN; c0 += 32) {
                umma::tmem_ld_wait();
                umma::tmem_ld16(ta + c0 + 16, a1);
                consume(a0, item_tile + c0, tile_full);
                consume(a0 + 8, item_tile + c0 + 8, tile_full);
                umma::tmem_ld_wait();
                if (c0 + 32 < TN) umma::tmem_ld16(ta + c0 + 32, a0);
                consume(a1, item_tile + c0 + 16, tile_full);
                consume(a1 + 8, item_tile + c0 + 24, tile_full);
            }
            umma::fence_before_sync();
            umma::mbar_arrive(&acc_empty[buf]);
        }
        if (sel && !flags[0]) {
            flush();
            // selection sort of the K keys, largest first (once per row and kernel)
            for (int p = 0; p < cnt; ++p) {
                unsigned long long best = lkr[p * TM];
                int bp = p;
                for (int j = p + 1; j < cnt; ++j) {
                    const unsigned long long k = lkr[j * TM];
                    if (k > best) { best = k; bp = j; }
                }
                lkr[bp * TM] = lkr[p * TM];
                lkr[p * TM] = best;
            }
            // the lists leave in SCALED units (s^2 u.i): the merge compares those, so that tiny embeddings whose
            // true scores underflow fp32 still rank correctly; the merge converts the final scores
            const size_t o = ((static_cast<size_t>(u0 + row) * S + sp) * N_GRP + grp) * K;
            for (int p = 0; p < K; ++p) {
                const bool ok = p < cnt;
                const unsigned long long k = ok ? lkr[p * TM] : 0ull;
                out_idx[o + p] = ok ? key_index(k) : -1;
                out_score[o + p] = ok ? key_score(k) : -INFINITY;
            }
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0 && flags[0]) atomicExch(err, 1);
    if (warp == 1) umma::tmem_dealloc(tmem, TMEM_COLS);
}

// L sorted lists per user (one per item range and read-out group) -> the user's top-K: highest score first,
// equal scores by lower item index - exactly what one list over the whole item axis would hold.
// Sixteen lanes per user, lane s owns list s: the lists are staged in shared memory with coalesced loads, then K
// rounds of a 16-lane arg-max over the list heads (shuffles), the winner's lane advances.  (One thread per user
// walking L list heads in global memory per round took 93 us for the 10 lists of an 8-way user shard.)
constexpr int MERGE_USERS = 8;                    // users per CTA (128 threads)

__global__ void __launch_bounds__(16 * MERGE_USERS)
topk_merge_kernel(const int32_t* __restrict__ part_idx, const float* __restrict__ part_score, int n_users, int L,
                  int K, int32_t* __restrict__ out_idx, float* __restrict__ out_score,
                  const unsigned int* __restrict__ word) {
    // dynamic shared memory, MERGE_USERS * L * K (index, score) pairs: sized for the call, not for L_MAX lists -
    // a static 32 KB array left 7 CTAs per SM and made the 2-list merge of an unsharded evaluation 3x slower
    extern __shared__ __align__(16) unsigned char merge_smem[];
    if (*reinterpret_cast<const volatile unsigned int*>(word + 1)) return;
    const int lu = threadIdx.x >> 4, l16 = threadIdx.x & 15;
    int32_t* const my_idx = reinterpret_cast<int32_t*>(merge_smem) + lu * L * K;
    float* const my_sc = reinterpret_cast<float*>(merge_smem) + (MERGE_USERS + lu) * L * K;
    const int u = blockIdx.x * MERGE_USERS + lu;
    if (u >= n_users) return;                             // whole 16-lane groups leave together
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);
    const float inv = 1.0f / operand_scale(word[0]);      // scores back to the caller's units: / s, twice
    const size_t base = static_cast<size_t>(u) * L * K;
    for (int i = l16; i < L * K; i += 16) {
        my_idx[i] = part_idx[base + i];
        my_sc[i] = part_score[base + i];
    }
    __syncwarp(gmask);
    int pos = 0;                                          // head of this lane's list
    for (int k = 0; k < K; ++k) {
        int idx = -1;
        float sc = -INFINITY;
        if (l16 < L && pos < K) {
            idx = my_idx[l16 * K + pos];
            sc = idx >= 0 ? my_sc[l16 * K + pos] : -INFINITY;
        }
        int b_idx = idx, b_lane = idx >= 0 ? l16 : -1;
        float b_sc = sc;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
            const int o_idx = __shfl_xor_sync(gmask, b_idx, o), o_lane = __shfl_xor_sync(gmask, b_lane, o);
            const float o_sc = __shfl_xor_sync(gmask, b_sc, o);
            const bool take = o_lane >= 0 && (b_lane < 0 || o_sc > b_sc || (o_sc == b_sc && o_idx < b_idx));
            if (take) { b_idx = o_idx; b_lane = o_lane; b_sc = o_sc; }
        }
        if (l16 == 0) {
            out_idx[static_cast<size_t>(u) * K + k] = b_lane >= 0 ? b_idx : -1;
            if (out_score) out_score[static_cast<size_t>(u) * K + k] = b_lane >= 0 ? b_sc * inv * inv : -INFINITY;
        }
        if (b_lane == l16) ++pos;
    }
}

// The same merge with one thread per user walking the list heads in global memory: the cheaper one for up to
// four lists (28 us against 81 us for the two lists per user of an unsharded evaluation), the slower one beyond
// (93 us for ten lists) - the launcher picks by L.
__global__ void __launch_bounds__(128)
topk_merge_serial_kernel(const int32_t* __restrict__ part_idx, const float* __restrict__ part_score, int n_users, int L,
                         int K, int32_t* __restrict__ out_idx, float* __restrict__ out_score,
                         const unsigned int* __restrict__ word) {
    if (*reinterpret_cast<const volatile unsigned int*>(word + 1)) return;
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n_users) return;
    const float inv = 1.0f / operand_scale(word[0]);
    int pos[4] = {0, 0, 0, 0};
    const size_t base = static_cast<size_t>(u) * L * K;
    for (int k = 0; k < K; ++k) {
        int best = -1, best_idx = -1;
        float best_sc = -INFINITY;
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            if (s >= L || pos[s] >= K) continue;
            const size_t o = base + static_cast<size_t>(s) * K + pos[s];
            const int idx = part_idx[o];
            if (idx < 0) continue;
            const float sc = part_score[o];
            if (best < 0 || sc > best_sc || (sc == best_sc && idx < best_idx)) { best = s; best_sc = sc; best_idx = idx; }
        }
        out_idx[static_cast<size_t>(u) * K + k] = best_idx;
        if (out_score) out_score[static_cast<size_t>(u) * K + k] = best >= 0 ? best_sc * inv * inv : -INFINITY;
#pragma unroll
        for (int s = 0; s < 4; ++s)
            if (s == best) ++pos[s];
    }
}

__device__ int g_eval_f16_error = 0;      // a tensor-core / copy completion barrier timed out
// [0] bits of max |x|, [1] fp32 kernel wanted (inf / NaN / dynamic range), [2] fallbacks taken, [3] bits of the
// smallest non-zero chunk maximum; [0], [1], [3] are per call
__device__ unsigned int g_eval_word[4] = {0, 0, 0, 0x7f800000u};

inline size_t align1k(size_t x) { return (x + 1023) / 1024 * 1024; }

// item ranges per user tile: fill the SMs at any user count without paying a unit's fixed cost more often than it
// helps.  That cost is not the user-tile load but the cold top-K lists: until a row has seen K good scores every
// score is a candidate and goes through the insertion path.  Measured on B200 (profiles/eval_shards.py): 3.65 us
// per item tile in steady state, ~95 us per unit on top = 26 tiles (the first version assumed 3 and cut an
// eighth of the users into 5-7 ranges: 1.18 ms for a 4-way shard where S = 1 takes 0.8).
constexpr int UNIT_OVERHEAD_TILES = 26;
inline int pick_splits(int u_tiles, int n_it, int sms) {
    int best_s = 1;
    long long best_cost = -1;
    for (int s = 1; s <= S_MAX && s <= n_it; ++s) {
        const long long waves = (static_cast<long long>(u_tiles) * s + sms - 1) / sms;
        const long long cost = waves * ((n_it + s - 1) / s + UNIT_OVERHEAD_TILES);
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
    p.part_idx = off; off += align1k(static_cast<size_t>(n_users) * L_MAX * K * 4);
    p.part_score = off; off += align1k(static_cast<size_t>(n_users) * L_MAX * K * 4);
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
    unsigned int v[4] = {0, 0, 0, 0}, zero = 0;
    if (cudaMemcpyFromSymbol(v, g_eval_word, sizeof(v)) != cudaSuccess) return KGAT_ERR_CUDA;
    if (v[2] != 0 && cudaMemcpyToSymbol(g_eval_word, &zero, sizeof(int), 2 * sizeof(int)) != cudaSuccess)
        return KGAT_ERR_CUDA;
    return static_cast<int>(v[2]);
}

int kgat_launch_score_topk_umma(const float* emb, int64_t emb_stride, int dim, const int32_t* user_ids, int n_users,
                                const int32_t* item_ids, int n_items, const int32_t* mask_ptr,
                                const int32_t* mask_items, int K, int32_t* topk_idx, float* topk_score,
                                int* only_if, cudaStream_t st);

// max |x| -> split (pre-scaled fp16 pairs) -> score + top-K per (user tile, item range, read-out group) -> merge;
// the 3xTF32 kernel runs behind iff an operand was inf / NaN (an empty launch otherwise)
int kgat_launch_score_topk_f16(const float* emb, int64_t emb_stride, int dim, const int32_t* user_ids, int n_users,
                               const int32_t* item_ids, int n_items, const int32_t* mask_ptr,
                               const int32_t* mask_items, int K, int32_t* topk_idx, float* topk_score,
                               void* workspace, size_t workspace_bytes, cudaStream_t st) {
    static PerDevice<int*> err_words;
    static PerDevice<unsigned int*> words;
    static PerDevice<size_t> smem_set;
    int*& err_word = err_words.get();
    unsigned int*& word = words.get();
    if (!err_word) KGAT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&err_word), g_eval_f16_error));
    if (!word) KGAT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&word), g_eval_word));
    const Plan p = make_plan(n_users, n_items, dim, K, num_sms());
    if (workspace_bytes < p.total) return KGAT_ERR_WORKSPACE;
    const size_t smem = smem_bytes(p.nkb, K);
    if (smem > smem_set.get()) {
        KGAT_CUDA(cudaFuncSetAttribute(score_topk_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set.get() = smem;
    }
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    KGAT_CUDA(cudaMemsetAsync(word, 0, 2 * sizeof(int), st));
    KGAT_CUDA(cudaMemsetAsync(word + 3, 0x7f, sizeof(int), st));        // 0x7f7f7f7f: a huge finite float
    const int64_t chunks = static_cast<int64_t>(p.u_tiles + p.i_tiles) * p.nkb * 1024;
    const unsigned pgrid = static_cast<unsigned>((chunks + 255) / 256);
    eval_absmax_kernel<<<pgrid, 256, 0, st>>>(emb, emb_stride, dim, user_ids, n_users, p.u_tiles, item_ids, n_items,
                                              p.i_tiles, p.nkb, word);
    KGAT_LAUNCH_CHECK();
    eval_split_kernel<<<pgrid, 256, 0, st>>>(emb, emb_stride, dim, user_ids, n_users, p.u_tiles, item_ids, n_items,
                                             p.i_tiles, p.nkb, ws + p.u_blob, ws + p.i_blob, word);
    KGAT_LAUNCH_CHECK();
    int32_t* pidx = reinterpret_cast<int32_t*>(ws + p.part_idx);
    float* pscore = reinterpret_cast<float*>(ws + p.part_score);
    score_topk_f16_kernel<<<static_cast<unsigned>(p.u_tiles) * p.S, THREADS, smem, st>>>(
        ws + p.u_blob, ws + p.i_blob, n_users, n_items, p.nkb, p.n_ks, p.S, mask_ptr, mask_items, K, pidx, pscore,
        err_word, word);
    KGAT_LAUNCH_CHECK();
    const int n_lists = p.S * N_GRP;
    if (n_lists <= 4) {
        topk_merge_serial_kernel<<<(n_users + 127) / 128, 128, 0, st>>>(pidx, pscore, n_users, n_lists, K, topk_idx,
                                                                      topk_score, word);
    } else {
        const size_t merge_bytes = static_cast<size_t>(2 * MERGE_USERS) * n_lists * K * sizeof(float);      // <= 32 KB
        topk_merge_kernel<<<(n_users + MERGE_USERS - 1) / MERGE_USERS, 16 * MERGE_USERS, merge_bytes, st>>>(
            pidx, pscore, n_users, n_lists, K, topk_idx, topk_score, word);
    }
    KGAT_LAUNCH_CHECK();
    return kgat_launch_score_topk_umma(emb, emb_stride, dim, user_ids, n_users, item_ids, n_items, mask_ptr, mask_items, K,
                                       topk_idx, topk_score, reinterpret_cast<int*>(word + 1), st);
}
