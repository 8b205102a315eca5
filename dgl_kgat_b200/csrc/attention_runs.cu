// (2') TransR attention logits through the (head, relation)-run shortcut, on the 5th-gen tensor cores.
//
//   pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r)  =  e_t . ( W_r tanh(e_h W_r + e_r) )     (reference models.py:135-144)
//
// The bracket depends on (h, r) only.  Inside a relation bucket the edges are head-sorted, so the edges that
// share (h, r) - a RUN - are contiguous: every user row of a CKG is one run, every item row one run per relation
// (kgat_runs_build).  Per run:  a = e_h W_r + e_r,  v = W_r tanh(a)   - two 64 x 64 mat-vecs, grouped into
// 128-run tiles of one relation and issued as tcgen05 split-fp16 GEMMs (attention_umma16.cu has the arithmetic);
// per edge what remains is a raw 64-wide dot  logit = e_t . v[run]  on the FFMA pipe.  Against the per-edge
// kernel (two projections per EDGE) the tensor-core work shrinks by E / n_runs x (1.9 x on the synthetic
// Amazon-book graph, whose KG part has uniformly random heads; far more on a real KG) and the producers
// gather one head row per run instead of a head and a tail row per edge.  Same math re-associated: the fp32
// difference to the reference order is ~1e-7.
//
// Roles of a persistent CTA (one per SM), tile = 128 runs of one relation:
//   producers (warps 0-3)   gather the 128 head rows, split x = hi + 2^-11 lo', store the K-major SWIZZLE_128B
//                           operand A1 (2-stage ring); on a relation change they also stage W_r twice:
//                           B1[n][k] = W[k][n] (for a = e_h W_r) and B2[i][j] = W[i][j] (for v = W_r x).
//   MMA warp (warp 12)      G1(t): acc1[slot] = A1 B1^T (12 MMAs, M = 128, N = 64, K = 16 each);
//                           G2(t): acc2[slot] = A2[slot] B2^T.  G1(t + 1) is issued before G2(t): the tanh of
//                           tile t runs on the read-out warps meanwhile.  Three tile slots in flight.
//   read-out groups (warps 4-7, 8-11; group = tile parity), thread = run row:
//                           step 1  x = tanh(2^-11 acc1 + e_r) -> split -> A2[slot] (swizzled fp16 hi / lo');
//                           step 2  v = 2^-11 acc2 -> this row of V[slot] (fp32, shares A2's memory);
//                           step 3  the warp walks the edges of ITS 32 runs (contiguous in bucket order):
//                                   16 lanes per edge read the tail row (128-bit coalesced) and the run's v
//                                   row from shared memory, 4 FMAs + a 16-lane reduction, logit -> CSR position.
// Range: the same limits and the same device-side protocol as attention_umma16.cu (|W_r| < 32 decided by
// att_range_kernel, non-finite logits reported through the fallback word, fp32-range kernel behind).
#include "common.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>

namespace kgat {
namespace {

constexpr int TILE = 128;
constexpr int D = 64;
constexpr int NS = 2;                 // A1 stages
constexpr int NTS = 3;                // tile slots: acc1 / acc2 in TMEM, A2 / V in shared memory
constexpr int N_PROD = 128;
constexpr int N_GRP = 2, N_EPI = 128;
constexpr int THREADS = N_PROD + N_GRP * N_EPI + 32;
constexpr int MMA_WARP = (N_PROD + N_GRP * N_EPI) / 32;
constexpr int MAX_REL = 1024;
constexpr uint32_t ARR_BYTES = TILE * D * 2;         // 16 KB: 128 rows x 64 fp16
constexpr uint32_t STAGE_BYTES = 2 * ARR_BYTES;      // hi, lo'
constexpr uint32_t W_ARR = 64 * 128;                 // 8 KB: 64 rows x 64 fp16
constexpr uint32_t TMEM_COLS = 512;
constexpr float LO_SCALE = 2048.0f, LO_INV = 1.0f / 2048.0f;
constexpr float TWO_LOG2E = 2.8853900817779268f;
constexpr int EPB = 16;               // edges per warp iteration of the edge walk

struct Smem {
    unsigned char B1[3 * W_ARR];                     // rows 0..63 2^11 W_hi^T, 64..127 W_lo'^T, 128..191 W_hi^T
    unsigned char B2[3 * W_ARR];                     // the same three arrays of W (not transposed)
    unsigned char A1[NS][STAGE_BYTES];
    unsigned char A2[NTS][STAGE_BYTES];              // A2 (fp16 hi / lo') until G2 has read it, then V (fp32 [128][64])
    float er[N_GRP][D];
    int prefix[MAX_REL + 1];                         // first tile of every relation
    int rstart[MAX_REL + 1];                         // rel_run_ptr, copied once
    uint64_t a1_full[NS], a1_empty[NS];
    uint64_t acc1_full[NTS], a2_full[NTS], acc2_full[NTS], slot_empty[NTS];
    uint64_t b_free, b_full, drain;
    uint32_t tmem_base;
    int t_range[2];
    volatile int abort_flag;
    volatile int skip;
};
static_assert(sizeof(Smem) + 1024 <= 227 * 1024, "shared-memory budget");

__device__ __forceinline__ float tanh_from_scaled(float y) {     // tanh(x) from y = 2 log2(e) x, ~1e-7 absolute
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(y)));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return copysignf(fmaf(-2.0f, r, 1.0f), y);
}

struct TileInfo { int rel, r0, n_r; };

__device__ __forceinline__ TileInfo tile_info(const int* prefix, const int* rstart, int tile, int& rel) {
    while (tile >= prefix[rel + 1]) ++rel;
    TileInfo t;
    t.rel = rel;
    t.r0 = rstart[rel] + (tile - prefix[rel]) * TILE;
    t.n_r = min(TILE, rstart[rel + 1] - t.r0);
    return t;
}

__device__ __forceinline__ uint32_t pack_h2(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    hi = pack_h2(h);
    lo = pack_h2(__floats2half2_rn((a - f.x) * LO_SCALE, (b - f.y) * LO_SCALE));
}

// position of input feature k inside the permuted 64-wide A1 row (chunk c holds k = 4c..4c+3, 32+4c..32+4c+3)
__device__ __forceinline__ int kpos(int k) { return ((k & 31) >> 2) * 8 + (k >> 5) * 4 + (k & 3); }

// byte offset of fp16 element (row, position p) in a SWIZZLE_128B tile with 128-byte rows
__device__ __forceinline__ uint32_t sw128_offset_h(int r, int p) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((p >> 3) ^ (r & 7)) & 7) << 4) + (p & 7) * 2);
}

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

__global__ void __launch_bounds__(THREADS, 1)
att_logits_runs_kernel(const int32_t* __restrict__ rel_run_ptr, int n_rel, const int32_t* __restrict__ run_ptr,
                       const int32_t* __restrict__ run_head, const int32_t* __restrict__ bk_run,
                       const int32_t* __restrict__ bk_tail, const int32_t* __restrict__ bk_pos,
                       const float* __restrict__ ent, const float* __restrict__ rel_emb,
                       const float* __restrict__ W_R, float* __restrict__ logits_csr,
                       float* __restrict__ v_out, int64_t v_base, const int32_t* __restrict__ rel_mode,
                       int* __restrict__ err, int* __restrict__ fallback) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) s.skip = *reinterpret_cast<volatile int*>(fallback);     // |W_r| >= 32: the fp32-range kernel runs
    __syncthreads();
    if (s.skip) return;

    // ---- one-time setup: tile prefix per relation, TMEM, barriers -------------------------------------
    if (warp == 0) {
        int base = 0;
        for (int r0 = 0; r0 < n_rel; r0 += 32) {
            const int r = r0 + lane;
            const int lo = r < n_rel ? rel_run_ptr[r] : 0, hi = r < n_rel ? rel_run_ptr[r + 1] : 0;
            // hybrid calls: only the relations marked 1 are ours, the per-edge kernel has the others
            const bool mine = !rel_mode || (r < n_rel && rel_mode[r] == 1);
            const int nt = mine ? (hi - lo + TILE - 1) / TILE : 0;
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
            umma::mbar_init(&s.a1_full[i], N_PROD);
            umma::mbar_init(&s.a1_empty[i], 1);
        }
        for (int i = 0; i < NTS; ++i) {
            umma::mbar_init(&s.acc1_full[i], 1);
            umma::mbar_init(&s.a2_full[i], N_EPI);
            umma::mbar_init(&s.acc2_full[i], 1);
            umma::mbar_init(&s.slot_empty[i], N_EPI);
        }
        umma::mbar_init(&s.b_free, 1);
        umma::mbar_init(&s.b_full, N_PROD);
        umma::mbar_init(&s.drain, 1);
        umma::fence_barrier_init();
        s.abort_flag = 0;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    // Contiguous tile ranges of equal COST, not equal tile count: a tile is 128 runs, but its edge walk is
    // proportional to its edges, and the runs of one relation can be 20 x longer than those of another (every item
    // row of the interaction relation vs a KG relation) - with equal counts the two CTAs that got the item
    // relation ran 13 x longer than the rest (measured: 4.5 ms instead of 0.35).  cost(boundary k) = RUN_COST *
    // runs before tile k + edges before tile k, monotone in k: two binary searches per CTA.
    // cost of everything before relation rel (relations without tiles - empty, or the per-edge kernel's in a hybrid
    // call - cost nothing); the table borrows A2[0], which no role touches before the ranges are known
    long long* const cbase = reinterpret_cast<long long*>(s.A2[0]);
    const long long RUN_COST = v_out ? 1000 : 3;           // split mode: no edge walk in this kernel
    if (tid == 0) {
        long long c = 0;
        for (int r = 0; r < n_rel; ++r) {
            cbase[r] = c;
            if (s.prefix[r + 1] > s.prefix[r])
                c += RUN_COST * (s.rstart[r + 1] - s.rstart[r]) +
                     static_cast<long long>(__ldg(run_ptr + s.rstart[r + 1]) - __ldg(run_ptr + s.rstart[r]));
        }
        cbase[n_rel] = c;
        const int n_tiles = s.prefix[n_rel];
        auto cost_at = [&](int k) {                        // cost of the tiles before tile k
            if (k >= n_tiles) return cbase[n_rel];
            int lo = 0, hi = n_rel - 1;                    // last relation with prefix[rel] <= k: the one that owns tile k
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s.prefix[mid] <= k) lo = mid; else hi = mid - 1;
            }
            const int r = s.rstart[lo] + (k - s.prefix[lo]) * TILE;
            return cbase[lo] + RUN_COST * (r - s.rstart[lo]) +
                   static_cast<long long>(__ldg(run_ptr + r) - __ldg(run_ptr + s.rstart[lo]));
        };
        const long long total = cost_at(n_tiles);
        for (int side = 0; side < 2; ++side) {
            const long long target = total * (blockIdx.x + side) / gridDim.x;
            int lo = 0, hi = n_tiles;                      // smallest k with cost_at(k) >= target
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (cost_at(mid) >= target) hi = mid; else lo = mid + 1;
            }
            s.t_range[side] = (blockIdx.x + side == gridDim.x) ? n_tiles : lo;
        }
    }
    __syncthreads();
    const int t0 = s.t_range[0], t1 = s.t_range[1];

    if (warp < 4) {
        // =========================================================== producers (head rows of the runs, W_r)
        const int wg = warp, c = lane & 7, rsub = lane >> 3;
        const int row_base = wg * 32 + rsub;                     // rows row_base + 4 j, j = 0..7
        uint32_t st_base[2];
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const int rr = par * 4 + rsub;
            st_base[par] = (wg * 4) * 1024 + rr * 128 + ((c ^ rr) << 4);
        }
        const float* __restrict__ ent_c = ent + c * 4;
        int ids[8];
        float4 v[16];
        int rel = 0, loaded_rel = -1, n_switch = 0;
        auto load_ids = [&](const TileInfo& t) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rl = row_base + 4 * j;
                ids[j] = rl < t.n_r ? __ldg(run_head + t.r0 + rl) : -1;
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
                const int it = tile - t0, stage = it % NS;
                if (loaded_rel != rel_cur) {
                    // W_r changes: the MMA warp says when nothing reads the old B1 / B2 any more
                    if (!umma::mbar_wait(&s.b_free, n_switch & 1)) s.abort_flag = 1;
                    const float* Wg = W_R + static_cast<size_t>(rel_cur) * D * D;
                    for (int idx = tid; idx < D * D; idx += N_PROD) {
                        const int k = idx >> 6, n = idx & 63;           // W[k][n]
                        const float w = __ldg(Wg + idx);
                        const __half hi = __float2half_rn(w);
                        const float hf = __half2float(hi);
                        const __half hs = __float2half_rn(hf * LO_SCALE), lo = __float2half_rn((w - hf) * LO_SCALE);
                        const int p = kpos(k);                          // a[n] = sum_k e_h[k] W[k][n]: row n, K position of k
                        *reinterpret_cast<__half*>(s.B1 + sw128_offset_h(n, p)) = hs;
                        *reinterpret_cast<__half*>(s.B1 + sw128_offset_h(64 + n, p)) = lo;
                        *reinterpret_cast<__half*>(s.B1 + sw128_offset_h(128 + n, p)) = hi;
                        // v[k] = sum_n W[k][n] x[n]: row k, K position n (natural order, as the read-out writes x)
                        *reinterpret_cast<__half*>(s.B2 + sw128_offset_h(k, n)) = hs;
                        *reinterpret_cast<__half*>(s.B2 + sw128_offset_h(64 + k, n)) = lo;
                        *reinterpret_cast<__half*>(s.B2 + sw128_offset_h(128 + k, n)) = hi;
                    }
                    umma::fence_proxy_async();
                    umma::mbar_arrive(&s.b_full);
                    loaded_rel = rel_cur;
                    ++n_switch;
                }
                if (!umma::mbar_wait(&s.a1_empty[stage], ((it / NS) & 1) ^ 1)) s.abort_flag = 1;
                unsigned char* base = s.A1[stage];
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
                umma::fence_proxy_async();
                umma::mbar_arrive(&s.a1_full[stage]);
                rel_cur = rel_nxt;
                if (tile + 1 < t1) {
                    issue();
                    if (tile + 2 < t1) {
                        ti = tile_info(s.prefix, s.rstart, tile + 2, rel);
                        rel_nxt = ti.rel;
                        load_ids(ti);
                    }
                }
            }
        }
    } else if (warp == MMA_WARP) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(TILE, D);
            const uint64_t dA0 = umma::make_smem_desc_sw128(umma::smem_u32(s.A1[0]));
            const uint32_t d_hi = static_cast<uint32_t>(dA0 >> 32);          // same upper word for every operand
            auto desc = [&](uint32_t lo) { return (static_cast<uint64_t>(d_hi) << 32) | lo; };
            const uint32_t a1_lo0 = static_cast<uint32_t>(dA0);
            const uint32_t a2_lo0 = static_cast<uint32_t>(umma::make_smem_desc_sw128(umma::smem_u32(s.A2[0])));
            const uint32_t b1_lo0 = static_cast<uint32_t>(umma::make_smem_desc_sw128(umma::smem_u32(s.B1)));
            const uint32_t b2_lo0 = static_cast<uint32_t>(umma::make_smem_desc_sw128(umma::smem_u32(s.B2)));
            // A x [2^11 W_hi | W_lo' | W_hi]-style triple into one accumulator: 2^11 (A . W)
            auto gemm = [&](uint32_t acc, uint32_t a_lo, uint32_t b_lo) {
#pragma unroll
                for (int ks = 0; ks < D / 16; ++ks) {
                    const uint32_t k16 = ks * 2;                  // K = 16 fp16 = 32 bytes inside the 128-byte row
                    const uint64_t dHi = desc(a_lo + k16);
                    mma_f16(acc, dHi, desc(b_lo + k16), idesc, ks != 0);                                  // hi x 2^11 W_hi
                    mma_f16(acc, dHi, desc(b_lo + (W_ARR >> 4) + k16), idesc, 1);                         // hi x W_lo'
                    mma_f16(acc, desc(a_lo + (ARR_BYTES >> 4) + k16), desc(b_lo + ((2 * W_ARR) >> 4) + k16), idesc, 1);   // lo' x W_hi
                }
            };
            auto do_g2 = [&](int it) {
                const int slot = it % NTS;
                if (!umma::mbar_wait(&s.a2_full[slot], (it / NTS) & 1)) s.abort_flag = 1;
                umma::fence_after_sync();
                gemm(tmem + slot * 128 + 64, a2_lo0 + slot * (STAGE_BYTES >> 4), b2_lo0);
                umma::mma_commit(&s.acc2_full[slot]);
            };
            int rel = 0, loaded_rel = -1, n_switch = 0;
            bool pending = false;
            for (int tile = t0; tile < t1 && !s.abort_flag; ++tile) {
                const int it = tile - t0, slot = it % NTS, stage = it % NS;
                const TileInfo ti = tile_info(s.prefix, s.rstart, tile, rel);
                if (ti.rel != loaded_rel) {
                    // relation change: finish what still reads the old W_r, then let the producers restage it
                    if (pending) { do_g2(it - 1); pending = false; }
                    if (n_switch > 0) {
                        umma::mma_commit(&s.drain);
                        if (!umma::mbar_wait(&s.drain, (n_switch - 1) & 1)) s.abort_flag = 1;
                    }
                    umma::mbar_arrive(&s.b_free);
                    if (!umma::mbar_wait(&s.b_full, n_switch & 1)) s.abort_flag = 1;
                    umma::fence_after_sync();
                    loaded_rel = ti.rel;
                    ++n_switch;
                }
                // G1(it)
                if (!umma::mbar_wait(&s.slot_empty[slot], ((it / NTS) & 1) ^ 1)) s.abort_flag = 1;
                if (!umma::mbar_wait(&s.a1_full[stage], (it / NS) & 1)) s.abort_flag = 1;
                umma::fence_after_sync();
                gemm(tmem + slot * 128, a1_lo0 + stage * (STAGE_BYTES >> 4), b1_lo0);
                umma::mma_commit(&s.a1_empty[stage]);
                umma::mma_commit(&s.acc1_full[slot]);
                // G2(it - 1): its tanh ran while G1(it) was being prepared
                if (pending) do_g2(it - 1);
                pending = true;
            }
            if (pending && !s.abort_flag) do_g2(t1 - t0 - 1);
        }
    } else {
        // =========================================================== read-out groups
        const int grp = (warp - 4) >> 2, q = warp & 3, row = q * 32 + lane, et = tid - N_PROD - grp * N_EPI;
        const int half = lane >> 4, l16 = lane & 15;
        int rel = 0, loaded_rel = -1;
        for (int tile = t0 + grp; tile < t1; tile += N_GRP) {
            const int it = tile - t0, slot = it % NTS;
            const TileInfo ti = tile_info(s.prefix, s.rstart, tile, rel);
            if (ti.rel != loaded_rel) {                               // group-uniform
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(N_EPI) : "memory");     // everyone is done with the old e_r
                if (et < D) s.er[grp][et] = __ldg(rel_emb + static_cast<size_t>(ti.rel) * D + et) * TWO_LOG2E;
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(N_EPI) : "memory");
                loaded_rel = ti.rel;
            }
            // ---- step 1: x = tanh(a), a = 2^-11 acc1 + e_r  ->  A2[slot] (this thread's row, natural K order)
            if (!umma::mbar_wait(&s.acc1_full[slot], (it / NTS) & 1)) s.abort_flag = 1;
            umma::fence_after_sync();
            if (s.abort_flag) break;
            const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + slot * 128;
            unsigned char* const a2 = s.A2[slot];
            const float* er = s.er[grp];
            const uint32_t row_off = (row >> 3) * 1024 + (row & 7) * 128;
            {
                float h[2][8];
                umma::tmem_ld8(ta, h[0]);
#pragma unroll
                for (int ch = 0; ch < D / 8; ++ch) {
                    umma::tmem_ld_wait();
                    if (ch + 1 < D / 8) umma::tmem_ld8(ta + (ch + 1) * 8, h[(ch + 1) & 1]);
                    float x[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        x[i] = tanh_from_scaled(fmaf(h[ch & 1][i], LO_INV * TWO_LOG2E, er[ch * 8 + i]));
                    uint4 hi, lo;
                    split_pair(x[0], x[1], hi.x, lo.x);
                    split_pair(x[2], x[3], hi.y, lo.y);
                    split_pair(x[4], x[5], hi.z, lo.z);
                    split_pair(x[6], x[7], hi.w, lo.w);
                    const uint32_t off = row_off + ((ch ^ (row & 7)) << 4);       // 16-byte chunk ch of the row
                    *reinterpret_cast<uint4*>(a2 + off) = hi;
                    *reinterpret_cast<uint4*>(a2 + ARR_BYTES + off) = lo;
                }
            }
            umma::fence_proxy_async();
            umma::mbar_arrive(&s.a2_full[slot]);
            // ---- step 2: v = 2^-11 acc2 -> V[slot][row] (fp32; chunk j of row r sits at chunk position j ^ (r & 15))
            if (!umma::mbar_wait(&s.acc2_full[slot], (it / NTS) & 1)) s.abort_flag = 1;
            umma::fence_after_sync();
            if (s.abort_flag) break;
            float* const V = reinterpret_cast<float*>(a2);            // G2 has consumed A2[slot]
            {
                float a[2][8];
                umma::tmem_ld8(ta + 64, a[0]);
#pragma unroll
                for (int ch = 0; ch < D / 8; ++ch) {
                    umma::tmem_ld_wait();
                    if (ch + 1 < D / 8) umma::tmem_ld8(ta + 64 + (ch + 1) * 8, a[(ch + 1) & 1]);
                    const float* c = a[ch & 1];
                    float* vr = V + row * D;
                    *reinterpret_cast<float4*>(vr + (((2 * ch) ^ (row & 15)) << 2)) =
                        make_float4(c[0] * LO_INV, c[1] * LO_INV, c[2] * LO_INV, c[3] * LO_INV);
                    *reinterpret_cast<float4*>(vr + (((2 * ch + 1) ^ (row & 15)) << 2)) =
                        make_float4(c[4] * LO_INV, c[5] * LO_INV, c[6] * LO_INV, c[7] * LO_INV);
                }
            }
            __syncwarp();
            const int w_first = ti.r0 + q * 32;
            const int w_last = min(w_first + 32, ti.r0 + ti.n_r);
            if (v_out) {
                // ---- step 3 (split mode): v rows -> global memory, coalesced (16 lanes per 256-byte row); the edge
                // dots run as a kernel of their own at full occupancy (edge_dot_kernel)
                for (int r = w_first + half; r < w_last; r += 2) {
                    const int rl = r - ti.r0;
                    const float4 vv = *reinterpret_cast<const float4*>(V + rl * D + ((l16 ^ (rl & 15)) << 2));
                    *reinterpret_cast<float4*>(v_out + static_cast<size_t>(r - v_base) * D + l16 * 4) = vv;
                }
            } else if (w_first < w_last) {
                // ---- step 3 (fused mode): the edges of this warp's 32 runs; 16 lanes per edge
                const int e_beg = __ldg(run_ptr + w_first), e_end = __ldg(run_ptr + w_last);
                // lane L < 16 holds the metadata of edge pbase + L of the current batch of 16
                int m_tail = 0, m_run = 0, m_pos = -1;
                auto load_meta = [&](int pbase) {
                    const int p = pbase + l16;
                    const bool ok = p < e_end;
                    m_tail = ok ? __ldg(bk_tail + p) : 0;
                    m_run = ok ? __ldg(bk_run + p) - w_first : 0;
                    m_pos = ok ? __ldg(bk_pos + p) : -1;
                };
                load_meta(e_beg);
                for (int pbase = e_beg; pbase < e_end; pbase += EPB) {
                    const int c_tail = m_tail, c_run = m_run, c_pos = m_pos;
                    if (pbase + EPB < e_end) load_meta(pbase + EPB);          // a batch ahead
                    float4 tv[EPB / 2];
                    int rr[EPB / 2];
#pragma unroll
                    for (int j = 0; j < EPB / 2; ++j) {                         // edge 2 j + half: all loads in flight
                        const int e = 2 * j + half;
                        const int tail = __shfl_sync(FULL_MASK, c_tail, e);
                        rr[j] = q * 32 + __shfl_sync(FULL_MASK, c_run, e);
                        tv[j] = ldg4(ent + static_cast<size_t>(tail) * D + l16 * 4);
                    }
#pragma unroll
                    for (int j = 0; j < EPB / 2; ++j) {
                        const int e = 2 * j + half;
                        const float4 vv = *reinterpret_cast<const float4*>(V + rr[j] * D + ((l16 ^ (rr[j] & 15)) << 2));
                        float p = tv[j].x * vv.x;
                        p = fmaf(tv[j].y, vv.y, p); p = fmaf(tv[j].z, vv.z, p); p = fmaf(tv[j].w, vv.w, p);
#pragma unroll
                        for (int o = 8; o > 0; o >>= 1) p += __shfl_xor_sync(FULL_MASK, p, o);
                        const int pos = __shfl_sync(FULL_MASK, c_pos, e);
                        if (l16 == 0 && pos >= 0) {
                            logits_csr[pos] = p;
                            // an entity value outside the fp16 range (head side) turns into inf / NaN here
                            if (!(fabsf(p) < INFINITY)) *reinterpret_cast<volatile int*>(fallback) = 1;
                        }
                    }
                }
            }
            __syncwarp();
            umma::fence_before_sync();
            umma::mbar_arrive(&s.slot_empty[slot]);                   // acc1, acc2 and A2 / V of the slot are free
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0 && s.abort_flag) atomicExch(err, 1);
    if (warp == MMA_WARP) umma::tmem_dealloc(tmem, TMEM_COLS);
}

// logit[bk_pos[p]] = e_tail(p) . v[bk_run[p]]  for the bucket positions [p_begin, p_end): 16 lanes per edge (128-bit
// coalesced row reads), 4 edges per lane group with all eight row loads in flight, 16-lane shuffle reduction.  The v
// rows of consecutive edges are the same row for a whole run, so their reads are L1 hits: taking the row from the
// previous edge's registers instead (tried, with 4 and with 8 edges per group) made the kernel SLOWER - 0.38 / 0.55
// against 0.33 ms at the Last-FM shape - the random tail rows out of L2 bound it, not the v rows.
// `rel_mode` (hybrid calls, else null): only edges whose relation is marked 1 belong to this kernel, and only their
// runs have a row in v (rows [v_base, v_base + v_count)).
constexpr int EDGES_PER_GROUP = 4;

__global__ void __launch_bounds__(256)
edge_dot_kernel(const int32_t* __restrict__ bk_run, const int32_t* __restrict__ bk_tail,
                const int32_t* __restrict__ bk_pos, int64_t p_begin, int64_t p_end, const float* __restrict__ ent,
                const float* __restrict__ v, int64_t v_base, int64_t v_count, const int32_t* __restrict__ rel_mode,
                const int32_t* __restrict__ csr_rel, float* __restrict__ logits_csr, int* __restrict__ fallback) {
    const int l16 = threadIdx.x & 15;
    const int64_t grp = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) >> 4;
    const int64_t p0 = p_begin + grp * EDGES_PER_GROUP;
    if (p0 >= p_end) return;                              // whole 16-lane groups leave together
    int tail[EDGES_PER_GROUP];
    int64_t vr[EDGES_PER_GROUP];
#pragma unroll
    for (int j = 0; j < EDGES_PER_GROUP; ++j) {
        const bool ok = p0 + j < p_end;
        tail[j] = ok ? __ldg(bk_tail + p0 + j) : 0;
        vr[j] = ok ? static_cast<int64_t>(__ldg(bk_run + p0 + j)) - v_base : -1;
        if (vr[j] >= v_count) vr[j] = -1;                 // hybrid: a run of the other kernel's relation
    }
    float4 a[EDGES_PER_GROUP], b[EDGES_PER_GROUP];
#pragma unroll
    for (int j = 0; j < EDGES_PER_GROUP; ++j) {
        a[j] = ldg4(ent + static_cast<size_t>(tail[j]) * D + l16 * 4);
        b[j] = ldg4(v + static_cast<size_t>(vr[j] < 0 ? 0 : vr[j]) * D + l16 * 4);
    }
    const unsigned gmask = 0xffffu << (threadIdx.x & 16);
#pragma unroll
    for (int j = 0; j < EDGES_PER_GROUP; ++j) {
        float p = a[j].x * b[j].x;
        p = fmaf(a[j].y, b[j].y, p); p = fmaf(a[j].z, b[j].z, p); p = fmaf(a[j].w, b[j].w, p);
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) p += __shfl_xor_sync(gmask, p, o);
        if (l16 == j && vr[j] >= 0) {
            const int pos = __ldg(bk_pos + p0 + j);
            if (rel_mode && __ldg(rel_mode + __ldg(csr_rel + pos)) != 1) continue;      // the per-edge kernel's edge
            logits_csr[pos] = p;
            if (!(fabsf(p) < INFINITY)) *reinterpret_cast<volatile int*>(fallback) = 1;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// run construction (bucket order: relation, head, edge id)
// ------------------------------------------------------------------------------------------------
__global__ void run_flags_kernel(const int32_t* __restrict__ bk_pos, const int32_t* __restrict__ bk_head,
                                 const int32_t* __restrict__ csr_rel, int64_t E, int32_t* __restrict__ flags) {
    const int64_t p = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (p >= E) return;
    int f = 1;
    if (p > 0) f = (bk_head[p] != bk_head[p - 1]) || (csr_rel[bk_pos[p]] != csr_rel[bk_pos[p - 1]]);
    flags[p] = f;
}

// block-wide inclusive scan of the flags in chunks of SCAN_TPB * SCAN_IPT, three kernels (sums, offsets, apply)
constexpr int SCAN_TPB = 256, SCAN_IPT = 16, SCAN_CHUNK = SCAN_TPB * SCAN_IPT;

__global__ void __launch_bounds__(SCAN_TPB) run_chunk_sums_kernel(const int32_t* __restrict__ flags, int64_t E,
                                                                  int32_t* __restrict__ chunk_sum) {
    __shared__ int s_w[SCAN_TPB / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
    int t = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i)
        if (base + i < E) t += flags[base + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(FULL_MASK, t, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0;
        for (int w = 0; w < SCAN_TPB / 32; ++w) a += s_w[w];
        chunk_sum[blockIdx.x] = a;
    }
}

// one CTA: exclusive prefix of the chunk sums (in place), total -> n_runs
__global__ void __launch_bounds__(1024) run_chunk_offsets_kernel(int32_t* __restrict__ chunk_sum, int n_chunks,
                                                                 int32_t* __restrict__ n_runs) {
    __shared__ int s_w[32];
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n_chunks; c0 += 1024) {
        const int c = c0 + threadIdx.x;
        const int v = c < n_chunks ? chunk_sum[c] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(FULL_MASK, incl, o);
            if ((threadIdx.x & 31) >= o) incl += u;
        }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = s_w[threadIdx.x], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(FULL_MASK, wi, o);
                if (threadIdx.x >= o) wi += u;
            }
            s_w[threadIdx.x] = wi - w;                                 // exclusive over warps
        }
        __syncthreads();
        const int excl = s_base + s_w[threadIdx.x >> 5] + incl - v;
        if (c < n_chunks) chunk_sum[c] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_base = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_runs[0] = s_base;
}

__global__ void __launch_bounds__(SCAN_TPB)
run_apply_kernel(const int32_t* __restrict__ flags, const int32_t* __restrict__ chunk_off, int64_t E,
                 const int32_t* __restrict__ bk_head, int32_t* __restrict__ bk_run, int32_t* __restrict__ run_ptr,
                 int32_t* __restrict__ run_head) {
    __shared__ int s_w[SCAN_TPB / 32];
    const int64_t base = static_cast<int64_t>(blockIdx.x) * SCAN_CHUNK + threadIdx.x * SCAN_IPT;
    int f[SCAN_IPT], t = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        f[i] = base + i < E ? flags[base + i] : 0;
        t += f[i];
    }
    int incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(FULL_MASK, incl, o);
        if ((threadIdx.x & 31) >= o) incl += u;
    }
    if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < (threadIdx.x >> 5); ++w) wbase += s_w[w];
    int run = chunk_off[blockIdx.x] + wbase + incl - t;            // runs started before this thread's first element
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        if (base + i < E) {
            run += f[i];
            bk_run[base + i] = run - 1;
            if (f[i]) {
                run_ptr[run - 1] = static_cast<int32_t>(base + i);
                run_head[run - 1] = bk_head[base + i];
            }
        }
    }
}

__global__ void run_finish_kernel(const int32_t* __restrict__ rel_ptr, int n_rel, int64_t E,
                                  const int32_t* __restrict__ bk_run, const int32_t* __restrict__ n_runs,
                                  int32_t* __restrict__ run_ptr, int32_t* __restrict__ rel_run_ptr) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    const int nr = n_runs[0];
    if (r <= n_rel) {
        const int p = rel_ptr[r];
        rel_run_ptr[r] = p < E ? bk_run[p] : nr;
    }
    if (r == 0) run_ptr[nr] = static_cast<int32_t>(E);
}

// rel_mode[r] = 1 when relation r is worth the run shortcut: at most a third as many runs as edges
__global__ void rel_mode_kernel(const int32_t* __restrict__ rel_ptr, const int32_t* __restrict__ rel_run_ptr, int n_rel,
                                int32_t* __restrict__ rel_mode) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rel) return;
    const int64_t edges = rel_ptr[r + 1] - rel_ptr[r], runs = rel_run_ptr[r + 1] - rel_run_ptr[r];
    rel_mode[r] = (edges > 0 && 3 * runs <= edges) ? 1 : 0;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" size_t kgat_runs_build_workspace_bytes(int64_t n_edges) {
    const int64_t E = n_edges > 0 ? n_edges : 1;
    const int64_t chunks = (E + SCAN_CHUNK - 1) / SCAN_CHUNK;
    return static_cast<size_t>((E + chunks + 64) * sizeof(int32_t));
}

extern "C" int kgat_runs_build(const kgat_graph_t* g, const kgat_runs_t* runs, void* workspace, size_t workspace_bytes,
                               kgat_stream_t stream) {
    KGAT_REQUIRE(g && runs && runs->bk_run && runs->run_ptr && runs->run_head && runs->rel_run_ptr && runs->n_runs);
    cudaStream_t st = as_stream(stream);
    const int64_t E = g->n_edges;
    const int R = static_cast<int>(g->n_rel);
    if (E == 0) {
        KGAT_CUDA(cudaMemsetAsync(runs->rel_run_ptr, 0, sizeof(int32_t) * (R + 1), st));
        KGAT_CUDA(cudaMemsetAsync(runs->n_runs, 0, sizeof(int32_t), st));
        KGAT_CUDA(cudaMemsetAsync(runs->run_ptr, 0, sizeof(int32_t), st));
        if (runs->rel_mode) KGAT_CUDA(cudaMemsetAsync(runs->rel_mode, 0, sizeof(int32_t) * R, st));
        return KGAT_OK;
    }
    KGAT_REQUIRE(workspace && workspace_bytes >= kgat_runs_build_workspace_bytes(E));
    int32_t* flags = static_cast<int32_t*>(workspace);
    int32_t* chunk = flags + E;
    const int chunks = static_cast<int>((E + SCAN_CHUNK - 1) / SCAN_CHUNK);
    run_flags_kernel<<<static_cast<unsigned>((E + 255) / 256), 256, 0, st>>>(g->bk_pos, g->bk_head, g->csr_rel, E, flags);
    KGAT_LAUNCH_CHECK();
    run_chunk_sums_kernel<<<chunks, SCAN_TPB, 0, st>>>(flags, E, chunk);
    KGAT_LAUNCH_CHECK();
    run_chunk_offsets_kernel<<<1, 1024, 0, st>>>(chunk, chunks, runs->n_runs);
    KGAT_LAUNCH_CHECK();
    run_apply_kernel<<<chunks, SCAN_TPB, 0, st>>>(flags, chunk, E, g->bk_head, runs->bk_run, runs->run_ptr, runs->run_head);
    KGAT_LAUNCH_CHECK();
    run_finish_kernel<<<(R + 1 + 255) / 256, 256, 0, st>>>(g->rel_ptr, R, E, runs->bk_run, runs->n_runs, runs->run_ptr,
                                                           runs->rel_run_ptr);
    KGAT_LAUNCH_CHECK();
    if (runs->rel_mode) {
        rel_mode_kernel<<<(R + 255) / 256, 256, 0, st>>>(g->rel_ptr, runs->rel_run_ptr, R, runs->rel_mode);
        KGAT_LAUNCH_CHECK();
    }
    return KGAT_OK;
}

int* kgat_umma_error_word();      // attention_umma.cu

// Launcher used by kgat_attention_logits_runs / _hybrid (attention.cu) for d = rd = 64.
// hybrid (rel_mode non-null): only the relations marked 1; v_rows then holds the runs [v_base, v_base + v_count)
// and the edge dots cover the bucket positions [p_begin, p_end) that span those relations.
int kgat_launch_att_logits_runs(const kgat_graph_t* g, const kgat_runs_t* runs, const float* entity,
                                const float* rel_emb, const float* W_R, float* logits_csr, float* v_rows,
                                int64_t v_base, int64_t v_count, int64_t p_begin, int64_t p_end,
                                const int32_t* rel_mode, int* fallback, cudaStream_t st) {
    static PerDeviceOnce configured;
    int* err_flag = kgat_umma_error_word();
    if (!err_flag) return KGAT_ERR_CUDA;
    if (configured.need()) {
        KGAT_CUDA(cudaFuncSetAttribute(att_logits_runs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
        configured.set();
    }
    // the run count lives on the device; the grid is sized for the worst case (n_runs <= E) and capped at the SMs
    const int64_t max_tiles = (rel_mode ? p_end - p_begin : g->n_edges) / TILE + g->n_rel;
    const int grid = static_cast<int>(max_tiles < num_sms() ? max_tiles : num_sms());
    att_logits_runs_kernel<<<grid, THREADS, sizeof(Smem), st>>>(
        runs->rel_run_ptr, static_cast<int>(g->n_rel), runs->run_ptr, runs->run_head, runs->bk_run, g->bk_tail, g->bk_pos,
        entity, rel_emb, W_R, logits_csr, v_rows, v_base, rel_mode, err_flag, fallback);
    KGAT_LAUNCH_CHECK();
    if (v_rows && p_end > p_begin) {
        const int64_t groups = (p_end - p_begin + EDGES_PER_GROUP - 1) / EDGES_PER_GROUP;
        edge_dot_kernel<<<static_cast<unsigned>((groups * 16 + 255) / 256), 256, 0, st>>>(
            runs->bk_run, g->bk_tail, g->bk_pos, p_begin, p_end, entity, v_rows, v_base, v_count, rel_mode, g->csr_rel,
            logits_csr, fallback);
        KGAT_LAUNCH_CHECK();
    }
    return KGAT_OK;
}
