// (2) TransR attention logits on the 5th-gen tensor cores, split-fp16 arithmetic.
//
//   pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r)                        (reference models.py:135-144)
//
// Same structure as attention_umma.cu (persistent CTA per SM, warp-specialised producer / MMA /
// epilogue roles over mbarriers, relation-bucketed 128-edge tiles), but the fp32 operands are split
// into two fp16 numbers instead of two TF32 numbers:
//     x = hi + lo,   hi = RN_fp16(x),   lo' = RN_fp16((x - hi) * 2^11)        (22+ significant bits)
//     2^11 x . w ~= hi.(2^11 w_hi) + hi.w_lo' + lo'.w_hi
// The 2^11 scale keeps lo' a NORMAL fp16 number whenever hi is (fp16 has 5 exponent bits), so the
// pair carries ~2^-23 relative error for 6e-5 <= |x| < 65504 and <= 1.5e-11 absolute below that; all
// three products carry the same 2^11 factor, so they accumulate into ONE fp32 TMEM accumulator and
// the epilogue multiplies by 2^-11.  Range limits of this kernel: |entity| < 65504, |W_r| < 32 (2^11 w_hi
// must stay finite in fp16).  They are enforced on the device without a host round trip (attention.cu):
// att_range_kernel raises fallback[0] when some |W_r| >= 32 and this kernel then does nothing; an entity
// value outside the range yields a non-finite logit, which the epilogue reports through the same word; the
// 3xTF32 kernel (fp32 range) launched right behind recomputes every logit when the word is set.
// Why fp16 and not TF32: a kind::f16 MMA moves half the operand bytes per multiply-add (the TF32
// kernel stores 128 KB and the tensor core fetches 224 KB of shared memory per tile, at 128 B/clk),
// a stage is 32 KB so six stages (three tiles) are in flight, and a single accumulator halves the
// TMEM read traffic of the epilogue (64 B/clk per SM), which was the critical path.
//
//   producers (warps 0-3: tail rows, warps 4-7: head rows)  one sub-tile = the 128 tail rows (or head
//            rows) of a tile, all 64 features: 8 consecutive lanes read one row (2 x 128 contiguous
//            bytes), convert, and write one conflict-free 16-byte chunk each of hi and lo' in the
//            K-major SWIZZLE_128B layout.  The K positions inside a row are permuted (chunk c holds
//            k = 4c..4c+3, 32+4c..32+4c+3) - the same permutation is applied to W_r, and a dot product
//            does not care.  A thread has exactly ONE batch of gathers in flight, issued right after
//            its fence + arrive: fence.proxy.async is a MEMBAR that waits for every outstanding load of
//            the thread, so a second batch in flight across it (as in attention_umma.cu) serialises the
//            two on the full gather latency.
//   MMA      (one lane of the last warp) per sub-tile, 4 K-steps of 16: A_hi x 2^11 W_hi, A_hi x W_lo',
//            A_lo' x W_hi (N = 64 each) into the same 64 TMEM columns; four accumulator buffers.
//   epilogue (warps 8-11, one TMEM lane quarter each; tile metadata prefetched one tile ahead, TMEM
//            loads double-buffered against the math)
//            logit = 2^-11 sum_c T[c] * tanh(2^-11 H[c] + e_r[c]) at 7 instructions per element, scattered
//            to the edge's CSR position.
#include "common.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>

namespace kgat {
namespace {

constexpr int TILE = 128;
constexpr int D = 64;
constexpr int NS = 6;                 // ring of sub-tile stages
constexpr int N_PROD = 256;
constexpr int N_GRP = 128;            // producer group: warps 0-3 tail rows, warps 4-7 head rows
constexpr int EPI_HALVES = 1;         // 2: a second set of epilogue warps takes output columns 32..63
constexpr int N_EPI = 128 * EPI_HALVES;
constexpr int THREADS = N_PROD + N_EPI + 32;
constexpr int MMA_WARP = (N_PROD + N_EPI) / 32;
constexpr int MAX_REL = 1024;
constexpr uint32_t ARR_BYTES = TILE * D * 2;         // 16 KB: 128 rows x 64 fp16
constexpr uint32_t STAGE_BYTES = 2 * ARR_BYTES;      // hi, lo'
constexpr int NACC = 4;                              // accumulator buffers: (T: 64 + H: 64) columns each
constexpr uint32_t TMEM_COLS = 512;
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;

struct Smem {
    unsigned char B[3 * (ARR_BYTES / 2)];            // rows 0..63 2^11 W_hi^T, 64..127 W_lo'^T, 128..191 W_hi^T
    unsigned char A[NS][STAGE_BYTES];
    float er[2][D];
    float part[2][TILE];                             // partial dot products of the upper column half
    int prefix[MAX_REL + 1];                         // first tile of every relation
    int rstart[MAX_REL + 1];                         // rel_ptr, copied once
    uint64_t full[NS], empty[NS], acc_full[NACC], acc_empty[NACC];
    uint32_t tmem_base;
    volatile int abort_flag;
    volatile int epi_abort;
};

static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "attention tile ring does not fit the 227 KB of an SM");

// tanh(x) from y = 2 log2(e) x:  tanh = sign(y) (1 - 2 / (1 + 2^|y|)).  ex2.approx.ftz / rcp.approx.ftz
// directly (the argument is never negative, so the denormal-range fix-ups __expf carries are dead
// weight): MUFU.EX2, FADD, MUFU.RCP, FFMA, LOP3.  ~1e-7 ABSOLUTE error, which is what the logit needs;
// 2^|y| = inf gives exactly 1.
constexpr float TWO_LOG2E = 2.8853900817779268f;
__device__ __forceinline__ float tanh_from_scaled(float y) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(y)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf(fmaf(-2.0f, r, 1.0f), y);
}

struct TileInfo { int rel, e0, n_e; };

// shared-memory only: the roles call this on their critical paths
__device__ __forceinline__ TileInfo tile_info(const int* prefix, const int* rstart, int tile, int& rel) {
    while (tile >= prefix[rel + 1]) ++rel;
    TileInfo t;
    t.rel = rel;
    t.e0 = rstart[rel] + (tile - prefix[rel]) * TILE;
    t.n_e = min(TILE, rstart[rel + 1] - t.e0);
    return t;
}

__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

// (a, b) -> packed hi pair and packed scaled-residual pair
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = pack_h2(h);
    lo = pack_h2(__floats2half2_rn((a - f.x) * LO_SCALE, (b - f.y) * LO_SCALE));
}

// position of input feature k inside the permuted 64-wide row
__device__ __forceinline__ int kpos(int k) { return ((k & 31) >> 2) * 8 + (k >> 5) * 4 + (k & 3); }

// byte offset of fp16 element (row, position p) in a SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128_offset_h(int r, int p) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((p >> 3) ^ (r & 7)) & 7) << 4) + (p & 7) * 2);
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

#ifdef KGAT_ATT_PROFILE
__device__ unsigned long long g_att_prof[160][12];
#define PROF_T0() const long long _t0 = clock64()
#define PROF_ADD(var) var += clock64() - _t0
#else
#define PROF_T0()
#define PROF_ADD(var)
#endif

__global__ void __launch_bounds__(THREADS, 1)
att_logits_umma16_kernel(const int32_t* __restrict__ rel_ptr, int n_rel, const int32_t* __restrict__ bk_head,
                         const int32_t* __restrict__ bk_tail, const int32_t* __restrict__ bk_pos,
                         const float* __restrict__ ent, const float* __restrict__ rel_emb,
                         const float* __restrict__ W_R, float* __restrict__ logits_csr,
                         int* __restrict__ err, int* __restrict__ fallback, const int32_t* __restrict__ rel_skip) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // Range protocol (attention.cu): fallback[0] was set by att_range_kernel when some |W_r| >= 32; a CTA
    // that sees it does nothing - the fp32-range kernel launched behind this one recomputes every logit.
    // One thread reads the word so the decision is CTA-uniform even if another CTA raises it meanwhile.
    if (tid == 0) s.epi_abort = *reinterpret_cast<volatile int*>(fallback);
    __syncthreads();
    if (s.epi_abort) return;

    // ---- one-time setup: tile prefix per relation, rel_ptr copy, TMEM, barriers -----------------
    if (warp == 0) {
        int base = 0;
        for (int r0 = 0; r0 < n_rel; r0 += 32) {
            const int r = r0 + lane;
            const int lo = r < n_rel ? rel_ptr[r] : 0, hi = r < n_rel ? rel_ptr[r + 1] : 0;
            // hybrid calls: relations marked 1 belong to the run kernel and get no tiles here
            const bool skip = rel_skip && r < n_rel && rel_skip[r] == 1;
            const int nt = skip ? 0 : (hi - lo + TILE - 1) / TILE;
            int incl = nt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            if (r < n_rel) {
                s.prefix[r + 1] = base + incl;
                s.rstart[r] = lo;
                s.rstart[r + 1] = hi;
            }
            base += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (lane == 0) s.prefix[0] = 0;
    }
    if (warp == MMA_WARP) umma::tmem_alloc(&s.tmem_base, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < NS; ++i) {
            umma::mbar_init(&s.full[i], N_GRP);
            umma::mbar_init(&s.empty[i], 1);
        }
        for (int i = 0; i < NACC; ++i) {
            umma::mbar_init(&s.acc_full[i], 1);
            umma::mbar_init(&s.acc_empty[i], N_EPI);
        }
        umma::fence_barrier_init();
        s.abort_flag = 0;
        s.epi_abort = 0;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    const int n_tiles = s.prefix[n_rel];
    const int t0 = static_cast<int>(static_cast<long long>(n_tiles) * blockIdx.x / gridDim.x);
    const int t1 = static_cast<int>(static_cast<long long>(n_tiles) * (blockIdx.x + 1) / gridDim.x);

    if (warp < 8) {
        // =========================================================== producers
        // group g fills every second stage of the ring (fill index f = 2 * it + g).  Slot j (0..7) of a
        // thread: row R_j = (warp & 3) * 32 + 4 * j + (lane >> 3); lane & 7 = 16-byte chunk c.
        const int g = warp >> 2, wg = warp & 3;
        const int c = lane & 7, rsub = lane >> 3;
        const int row_base = wg * 32 + rsub;
        uint32_t st_base[2];
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const int rr = par * 4 + rsub;                          // row inside its 8-row group
            st_base[par] = (wg * 4) * 1024 + rr * 128 + ((c ^ rr) << 4);
        }
        const int32_t* __restrict__ id_src = g ? bk_head : bk_tail;
        const float* __restrict__ ent_c = ent + c * 4;
        int ids[8];                 // node ids of the NEXT batch to issue (-1: padding row)
        float4 v[16];               // the one batch in flight
        int rel = 0, loaded_rel = -1;
        int stage = g;
        uint32_t phase = 0;
        long long w_empty = 0, w_total = clock64();

        auto load_ids = [&](const TileInfo& t) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rl = row_base + 4 * j;
                ids[j] = rl < t.n_e ? __ldg(id_src + t.e0 + rl) : -1;
            }
        };
        auto issue = [&]() {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float* p = ent_c + static_cast<size_t>(ids[j] < 0 ? 0 : ids[j]) * D;
                v[2 * j] = ids[j] >= 0 ? ldg4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[2 * j + 1] = ids[j] >= 0 ? ldg4(p + 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        };

        if (t0 < t1) {
            TileInfo ti = tile_info(s.prefix, s.rstart, t0, rel);
            int rel_cur = ti.rel, rel_nxt = ti.rel;
            load_ids(ti);
            issue();
            if (t0 + 1 < t1) {
                ti = tile_info(s.prefix, s.rstart, t0 + 1, rel);
                rel_nxt = ti.rel;
                load_ids(ti);
            }
            for (int tile = t0; tile < t1 && !s.abort_flag; ++tile) {
                const int f = 2 * (tile - t0) + g;
                if (g == 0 && loaded_rel != rel_cur) {
                    // W_r changes: every MMA issued so far (last: the head sub-tile of the previous tile,
                    // fill f - 1) must be done reading B; they complete in order
                    if (f > 0 && !umma::mbar_wait(&s.empty[(f - 1) % NS], ((f - 1) / NS) & 1)) s.abort_flag = 1;
                    const float* Wg = W_R + static_cast<size_t>(rel_cur) * D * D;
                    for (int idx = tid; idx < D * D; idx += N_GRP) {
                        const int k = idx >> 6, n = idx & 63;
                        const float w = __ldg(Wg + idx);
                        const __half hi = __float2half_rn(w);
                        const float hf = __half2float(hi);
                        const int p = kpos(k);
                        *reinterpret_cast<__half*>(s.B + sw128_offset_h(n, p)) = __float2half_rn(hf * LO_SCALE);
                        *reinterpret_cast<__half*>(s.B + sw128_offset_h(64 + n, p)) =
                            __float2half_rn((w - hf) * LO_SCALE);
                        *reinterpret_cast<__half*>(s.B + sw128_offset_h(128 + n, p)) = hi;
                    }
                    loaded_rel = rel_cur;
                }
                {
                    PROF_T0();
                    if (!umma::mbar_wait(&s.empty[stage], phase ^ 1)) s.abort_flag = 1;
                    PROF_ADD(w_empty);
                }
                unsigned char* base = s.A[stage];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t off = st_base[j & 1] + (j >> 1) * 1024;
                    uint4 hi, lo;
                    split_pair(v[2 * j].x, v[2 * j].y, hi.x, lo.x);
                    split_pair(v[2 * j].z, v[2 * j].w, hi.y, lo.y);
                    split_pair(v[2 * j + 1].x, v[2 * j + 1].y, hi.z, lo.z);
                    split_pair(v[2 * j + 1].z, v[2 * j + 1].w, hi.w, lo.w);
                    *reinterpret_cast<uint4*>(base + off) = hi;
                    *reinterpret_cast<uint4*>(base + ARR_BYTES + off) = lo;
                }
                umma::fence_proxy_async();                           // nothing of this thread is in flight here
                umma::mbar_arrive(&s.full[stage]);
                stage += 2;
                if (stage >= NS) { stage -= NS; phase ^= 1; }
                rel_cur = rel_nxt;
                if (tile + 1 < t1) {
                    issue();                                         // the next tile's rows
                    if (tile + 2 < t1) {
                        ti = tile_info(s.prefix, s.rstart, tile + 2, rel);
                        rel_nxt = ti.rel;
                        load_ids(ti);
                    }
                }
            }
        }
#ifdef KGAT_ATT_PROFILE
        if (tid == 0) { g_att_prof[blockIdx.x][0] = clock64() - w_total; g_att_prof[blockIdx.x][1] = w_empty; }
#endif
    } else if (warp == MMA_WARP) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            // The issuing thread is a single in-order instruction stream: keep it short.  A descriptor's
            // start-address field is its low 14 bits (address >> 4), so stepping through stages, the hi / lo'
            // arrays and the K-steps is one 32-bit add on a precomputed descriptor.
            const uint32_t idesc = make_idesc_f16(TILE, D);
            const uint64_t dA0 = umma::make_smem_desc_sw128(umma::smem_u32(s.A[0]));
            const uint64_t dB0 = umma::make_smem_desc_sw128(umma::smem_u32(s.B));
            const uint32_t d_hi = static_cast<uint32_t>(dA0 >> 32);          // same upper word for every operand
            const uint32_t a_lo0 = static_cast<uint32_t>(dA0), b_lo0 = static_cast<uint32_t>(dB0);
            auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(d_hi) << 32) | lo; };
            int stage = 0;
            uint32_t phase = 0;
            long long w_full = 0, w_accE = 0, w_total = clock64();
            for (int tile = t0; tile < t1 && !s.abort_flag; ++tile) {
                const int it = tile - t0, buf = it % NACC;
                {
                    PROF_T0();
                    if (!umma::mbar_wait(&s.acc_empty[buf], ((it / NACC) & 1) ^ 1)) s.abort_flag = 1;
                    PROF_ADD(w_accE);
                }
                umma::fence_after_sync();
#pragma unroll
                for (int o = 0; o < 2; ++o) {                         // tail rows, then head rows
                    {
                        PROF_T0();
                        if (!umma::mbar_wait(&s.full[stage], phase)) s.abort_flag = 1;
                        PROF_ADD(w_full);
                    }
                    umma::fence_after_sync();
                    const uint32_t a_lo = a_lo0 + stage * (STAGE_BYTES >> 4);
                    const uint32_t acc = tmem + buf * 128 + o * 64;
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t k16 = ks * 2;                  // K = 16 fp16 = 32 bytes inside the 128-byte row
                        const uint64_t dHi = desc(a_lo + k16);
                        mma_f16(acc, dHi, desc(b_lo0 + k16), idesc, ks != 0);                         // x 2^11 W_hi
                        mma_f16(acc, dHi, desc(b_lo0 + ((64 * 128) >> 4) + k16), idesc, 1);           // x W_lo'
                        mma_f16(acc, desc(a_lo + (ARR_BYTES >> 4) + k16), desc(b_lo0 + ((128 * 128) >> 4) + k16),
                                idesc, 1);                                                             // lo' x W_hi
                    }
                    umma::mma_commit(&s.empty[stage]);
                    if (++stage == NS) { stage = 0; phase ^= 1; }
                }
                umma::mma_commit(&s.acc_full[buf]);
            }
#ifdef KGAT_ATT_PROFILE
            g_att_prof[blockIdx.x][2] = clock64() - w_total; g_att_prof[blockIdx.x][3] = w_full;
            g_att_prof[blockIdx.x][4] = w_accE;
#endif
        }
    } else {
        // =========================================================== epilogue
        const int q = warp & 3;
        const int upper = EPI_HALVES == 2 ? (warp - 8) >> 2 : 0;      // 1: output columns 32..63
        constexpr int COLS = D / EPI_HALVES;
        const int row = q * 32 + lane;
        const int et = tid - N_PROD;
        int rel = 0;
        long long w_accF = 0, w_bar = 0, w_total = clock64();
        // metadata of the next tile is fetched while the current one is computed
        int pos_n = -1;
        float er_n = 0.f;
        if (t0 < t1) {
            const TileInfo t = tile_info(s.prefix, s.rstart, t0, rel);
            if (!upper && row < t.n_e) pos_n = __ldg(bk_pos + t.e0 + row);
            if (et < D) er_n = __ldg(rel_emb + static_cast<size_t>(t.rel) * D + et);
        }
        for (int tile = t0; tile < t1; ++tile) {
            const int it = tile - t0, buf = it % NACC, eb = it & 1;
            const int pos = pos_n;
            if (et < D) s.er[eb][et] = er_n * TWO_LOG2E;              // e_r pre-scaled for tanh_from_scaled
            pos_n = -1;
            if (tile + 1 < t1) {
                const TileInfo t = tile_info(s.prefix, s.rstart, tile + 1, rel);
                if (!upper && row < t.n_e) pos_n = __ldg(bk_pos + t.e0 + row);
                if (et < D) er_n = __ldg(rel_emb + static_cast<size_t>(t.rel) * D + et);
            }
            {
                PROF_T0();
                if (!umma::mbar_wait(&s.acc_full[buf], (it / NACC) & 1)) { s.abort_flag = 1; s.epi_abort = 1; }
                PROF_ADD(w_accF);
            }
            if (et == 0 && s.abort_flag) s.epi_abort = 1;
            umma::fence_after_sync();
            {
                PROF_T0();
                asm volatile("bar.sync 1, %0;" ::"n"(N_EPI) : "memory");  // er[eb] visible, abort view agreed
                PROF_ADD(w_bar);
            }
            if (s.epi_abort) break;
            const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * 128 + upper * COLS;
            const float* er = s.er[eb] + upper * COLS;
            float p = 0.f;
            constexpr int CW = 8;                                     // chunk k + 1 loads while chunk k is computed
            float a[2][CW], h[2][CW];
            umma::tmem_ld8(ta, a[0]);
            umma::tmem_ld8(ta + 64, h[0]);
#pragma unroll
            for (int ch = 0; ch < COLS / CW; ++ch) {
                umma::tmem_ld_wait();
                if (ch + 1 < COLS / CW) {
                    umma::tmem_ld8(ta + (ch + 1) * CW, a[(ch + 1) & 1]);
                    umma::tmem_ld8(ta + 64 + (ch + 1) * CW, h[(ch + 1) & 1]);
                }
#pragma unroll
                for (int i = 0; i < CW; ++i)
                    p = fmaf(a[ch & 1][i], tanh_from_scaled(fmaf(h[ch & 1][i], LO_INV * TWO_LOG2E, er[ch * CW + i])), p);
            }
            p *= LO_INV;
            umma::fence_before_sync();
            umma::mbar_arrive(&s.acc_empty[buf]);                     // TMEM buffer may be overwritten
            if constexpr (EPI_HALVES == 2) {
                if (upper) {
                    s.part[eb][row] = p;
                    asm volatile("bar.arrive 2, %0;" ::"n"(N_EPI) : "memory");
                } else {
                    asm volatile("bar.sync 2, %0;" ::"n"(N_EPI) : "memory");
                    if (pos >= 0) {
                        const float v = p + s.part[eb][row];
                        logits_csr[pos] = v;
                        if (!(fabsf(v) < INFINITY)) *reinterpret_cast<volatile int*>(fallback) = 1;
                    }
                }
            } else {
                if (pos >= 0) {
                    logits_csr[pos] = p;
                    // an entity value outside the fp16 range turns into inf/NaN here: ask for the fp32-range kernel
                    if (!(fabsf(p) < INFINITY)) *reinterpret_cast<volatile int*>(fallback) = 1;
                }
            }
        }
#ifdef KGAT_ATT_PROFILE
        if (et == 0) { g_att_prof[blockIdx.x][5] = clock64() - w_total; g_att_prof[blockIdx.x][6] = w_accF;
                       g_att_prof[blockIdx.x][7] = t1 - t0; g_att_prof[blockIdx.x][8] = w_bar;
                       g_att_prof[blockIdx.x][9] = 0; }
#endif
    }

    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0 && s.abort_flag) atomicExch(err, 1);
    if (warp == MMA_WARP) umma::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace
}  // namespace kgat

using namespace kgat;

#ifdef KGAT_ATT_PROFILE
// debug build only: per-CTA role timings {prod total, prod wait empty, mma total, mma wait full,
// mma wait acc_empty, epi total, epi wait acc_full, tiles}
extern "C" KGAT_API int kgat_debug_att_profile(unsigned long long* out) {
    return cudaMemcpyFromSymbol(out, g_att_prof, sizeof(unsigned long long) * 160 * 12) == cudaSuccess ? 0 : -4;
}
#endif

int* kgat_umma_error_word();      // attention_umma.cu

// Launcher used by kgat_attention_logits for d = rd = 64 (declared in attention.cu).
int kgat_launch_att_logits_umma16(const kgat_graph_t* g, const float* entity, const float* rel_emb,
                                  const float* W_R, float* logits_csr, int* fallback, const int32_t* rel_skip,
                                  cudaStream_t st) {
    static PerDeviceOnce configured;
    int* err_flag = kgat_umma_error_word();
    if (!err_flag) return KGAT_ERR_CUDA;
    if (configured.need()) {
        KGAT_CUDA(cudaFuncSetAttribute(att_logits_umma16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
        configured.set();
    }
    const int64_t max_tiles = g->n_edges / TILE + g->n_rel;
    const int grid = static_cast<int>(max_tiles < num_sms() ? max_tiles : num_sms());
    att_logits_umma16_kernel<<<grid, THREADS, sizeof(Smem), st>>>(
        g->rel_ptr, static_cast<int>(g->n_rel), g->bk_head, g->bk_tail, g->bk_pos, entity, rel_emb, W_R, logits_csr,
        err_flag, fallback, rel_skip);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
