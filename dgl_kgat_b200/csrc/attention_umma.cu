// (2) TransR attention logits on the 5th-gen tensor cores: relation-grouped projection as a
// tcgen05 GEMM in 3xTF32, fused with the per-edge gather (prologue) and the tanh / dot epilogue.
//
//   pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r)                        (reference models.py:135-144)
//
// One persistent CTA per SM walks a contiguous range of 128-edge tiles of the relation buckets.
// Warp-specialised, three roles connected by mbarriers:
//   producers (warps 0-7)  gather the 128 tail rows and 128 head rows of the tile from the entity table
//            (128-bit loads; 8 consecutive lanes read one row's contiguous 128 bytes, a warp instruction
//            covers 4 rows), split every value x = hi + lo (hi = RN-TF32(x), lo exact) and store both
//            halves in shared memory in the K-major SWIZZLE_128B layout (umma.cuh): the same 8 lanes
//            write one conflict-free 128-byte row.  A stage holds one K-half (32 of the 64 input
//            features = one 128-byte row per edge) of a tile: 4 arrays x 16 KB; two stages ring.  Loads for the next stage are in
//            flight while the current one is split and stored.  W_r is staged as B = [W_hi | W_lo]
//            (N = 128) once per relation.
//   MMA      (warp 12, one lane) per stage: T_hi x W_hi, T_hi x W_lo, T_lo x W_hi (N=64 each) accumulate
//            into the SAME 64 TMEM columns, and the same for the head rows; fp32 accumulation in TMEM,
//            two accumulator buffers (2 x 128 columns).  tcgen05.commit releases the shared-memory
//            stage and, after the second K-half, publishes the accumulator.
//   epilogue (warps 8-11, one TMEM lane quarter each) tcgen05.ld,
//            logit = sum_c t_r[c] * tanh(h_r[c] + e_r[c]), scattered to the edge's CSR position; runs
//            while the tensor core works on the next tile.
// Neither projection ever exists in HBM.  3xTF32 keeps the result inside the 1e-5 relative bar
// (a single TF32 pass is ~4e-4).
#include "common.cuh"
#include "umma.cuh"

#include <cstdlib>

namespace kgat {
namespace {

constexpr int TILE = 128;
constexpr int D = 64;                 // entity and relation width of this specialisation
constexpr int KH = 32;                // K-half held by one stage
constexpr int N_PROD = 256;           // producer threads (warps 0-7: 0-3 tail rows, 4-7 head rows)
constexpr int N_EPI = 128;            // epilogue threads (warps 8-11, TMEM lane quarter = warp % 4)
constexpr int THREADS = N_PROD + N_EPI + 32;
constexpr int MMA_WARP = 12;
constexpr int MAX_REL = 1024;
constexpr uint32_t ARR_BYTES = TILE * KH * 4;        // 16 KB: one operand array of one stage
constexpr uint32_t STAGE_BYTES = 4 * ARR_BYTES;      // T_hi, T_lo, H_hi, H_lo
constexpr uint32_t B_HALF_BYTES = 128 * KH * 4;      // 16 KB: [W_hi | W_lo] rows for one K-half
constexpr uint32_t TMEM_COLS = 512;

struct Smem {
    unsigned char B[2][B_HALF_BYTES];                // per K-half: rows 0..63 W_hi^T, rows 64..127 W_lo^T
    unsigned char A[2][STAGE_BYTES];
    float er[2][D];
    int prefix[MAX_REL + 1];
    uint64_t full[2], empty[2], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    volatile int abort_flag;                         // a bounded barrier wait timed out somewhere
    volatile int epi_abort;                          // the epilogue group's agreed view of it
};

// tanh with ~1e-7 ABSOLUTE error (what the logit needs): 1 - 2 / (1 + exp(2|x|)), sign restored
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(2.0f * fabsf(x));
    return copysignf(1.0f - __fdividef(2.0f, 1.0f + e), x);
}

struct TileInfo { int rel, e0, n_e; };

__device__ __forceinline__ TileInfo tile_info(const int* prefix, const int32_t* __restrict__ rel_ptr, int tile,
                                              int& rel) {
    while (tile >= prefix[rel + 1]) ++rel;
    TileInfo t;
    t.rel = rel;
    t.e0 = __ldg(rel_ptr + rel) + (tile - prefix[rel]) * TILE;
    t.n_e = min(TILE, __ldg(rel_ptr + rel + 1) - t.e0);
    return t;
}

template <int EPI_W, int WIDE>
__global__ void __launch_bounds__(THREADS, 1)
att_logits_umma_kernel(const int32_t* __restrict__ rel_ptr, int n_rel, const int32_t* __restrict__ bk_head,
                       const int32_t* __restrict__ bk_tail, const int32_t* __restrict__ bk_pos,
                       const float* __restrict__ ent, const float* __restrict__ rel_emb,
                       const float* __restrict__ W_R, float* __restrict__ logits_csr,
                       int* __restrict__ err, int* __restrict__ only_if) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // Launched behind the split-fp16 kernel as its fp32-range fallback: only_if[0] != 0 means that kernel
    // declined (|W_r| >= 32) or produced a non-finite logit (an operand left the fp16 range); otherwise
    // there is nothing to do.  The word is final here (stream order), so the exit is grid-uniform.
    if (only_if) {
        if (*reinterpret_cast<volatile int*>(only_if) == 0) return;
        if (blockIdx.x == 0 && tid == 0) atomicAdd(only_if + 1, 1);     // fallbacks taken (kgat_attention_fallbacks)
    }

    // ---- one-time setup: tile prefix per relation, TMEM, barriers -------------------------------
    if (warp == 0) {
        int base = 0;
        for (int r0 = 0; r0 < n_rel; r0 += 32) {
            const int r = r0 + lane;
            const int nt = r < n_rel ? (rel_ptr[r + 1] - rel_ptr[r] + TILE - 1) / TILE : 0;
            int incl = nt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL_MASK, incl, o);
                if (lane >= o) incl += v;
            }
            if (r < n_rel) s.prefix[r + 1] = base + incl;
            base += __shfl_sync(FULL_MASK, incl, 31);
        }
        if (lane == 0) s.prefix[0] = 0;
    }
    if (warp == MMA_WARP) umma::tmem_alloc(&s.tmem_base, TMEM_COLS);
    if (tid == 32) {
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&s.full[i], N_PROD);
            umma::mbar_init(&s.empty[i], 1);
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
        // Per stage (one K-half) a thread moves 8 x 16 bytes: slot j -> row R_j = (warp & 3) * 32 + 4 * j +
        // (lane >> 3), 16-byte chunk c = lane & 7.  Eight consecutive lanes cover one row's 128 bytes, in
        // global memory and (chunk-permuted) in shared memory.
        const int c = lane & 7, rsub = lane >> 3;
        const int32_t* __restrict__ id_src = warp < 4 ? bk_tail : bk_head;
        const int row_base = (warp & 3) * 32 + rsub;                     // R_j = row_base + 4 * j
        uint32_t st_base[2];                                             // by parity of j: row-in-group = (j&1)*4 + rsub
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const int rr = par * 4 + rsub;
            st_base[par] = (warp < 4 ? 0u : 2u * ARR_BYTES) + ((warp & 3) * 4) * 1024 + rr * 128 + ((c ^ rr) << 4);
        }
        const float* __restrict__ ent_c = ent + c * 4;
        int ids[8], ids_next[8];    // node ids of this thread's 8 rows (-1: padding row)
        float4 va[8], vb[8];        // in-flight gathers: K-half 0 / K-half 1
        int rel = 0, loaded_rel = -1;

        auto load_ids = [&](const TileInfo& t, int* dst) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int rl = row_base + 4 * j;
                dst[j] = rl < t.n_e ? __ldg(id_src + t.e0 + rl) : -1;
            }
        };
        auto issue = [&](float4* v, const int* id, int h) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = id[j] >= 0 ? ldg4(ent_c + static_cast<size_t>(id[j]) * D + h * KH)
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
        };
        auto store = [&](const float4* v, int stage) {
            unsigned char* base = s.A[stage];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint32_t off = st_base[j & 1] + (j >> 1) * 1024;
                float4 hi, lo;
                umma::split_tf32(v[j].x, hi.x, lo.x);
                umma::split_tf32(v[j].y, hi.y, lo.y);
                umma::split_tf32(v[j].z, hi.z, lo.z);
                umma::split_tf32(v[j].w, hi.w, lo.w);
                *reinterpret_cast<float4*>(base + off) = hi;
                *reinterpret_cast<float4*>(base + ARR_BYTES + off) = lo;
            }
        };

        if (t0 < t1) {
            TileInfo cur = tile_info(s.prefix, rel_ptr, t0, rel);
            load_ids(cur, ids);
            issue(va, ids, 0);
            for (int tile = t0; tile < t1 && !s.abort_flag; ++tile) {
                const int it = tile - t0;
                const bool more = tile + 1 < t1;
                TileInfo nxt = cur;
                if (more) {                                          // ids of the next tile: a full tile ahead
                    nxt = tile_info(s.prefix, rel_ptr, tile + 1, rel);
                    load_ids(nxt, ids_next);
                }
                issue(vb, ids, 1);                                   // K-half 1 of this tile
                // ---- stage 0 <- K-half 0
                if (!umma::mbar_wait(&s.empty[0], (it & 1) ^ 1)) s.abort_flag = 1;
                if (loaded_rel != cur.rel) {
                    // W_r changes: every MMA issued so far must be done reading B (they complete in order)
                    if (it > 0 && !umma::mbar_wait(&s.empty[1], (it - 1) & 1)) s.abort_flag = 1;
                    const float* Wg = W_R + static_cast<size_t>(cur.rel) * D * D;
                    for (int idx = tid; idx < D * D; idx += N_PROD) {
                        const int k = idx >> 6, n = idx & 63;
                        float hi, lo;
                        umma::split_tf32(__ldg(Wg + idx), hi, lo);
                        *reinterpret_cast<float*>(s.B[k >> 5] + umma::sw128_offset(n, k & 31)) = hi;
                        *reinterpret_cast<float*>(s.B[k >> 5] + umma::sw128_offset(64 + n, k & 31)) = lo;
                    }
                    loaded_rel = cur.rel;
                }
                store(va, 0);
                umma::fence_proxy_async();
                umma::mbar_arrive(&s.full[0]);
                if (more) issue(va, ids_next, 0);                    // K-half 0 of the next tile
                // ---- stage 1 <- K-half 1
                if (!umma::mbar_wait(&s.empty[1], (it & 1) ^ 1)) s.abort_flag = 1;
                store(vb, 1);
                umma::fence_proxy_async();
                umma::mbar_arrive(&s.full[1]);
#pragma unroll
                for (int j = 0; j < 8; ++j) ids[j] = ids_next[j];
                cur = nxt;
            }
        }
    } else if (warp == MMA_WARP) {
        // =========================================================== MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma::make_idesc_tf32(TILE, D), idesc128 = umma::make_idesc_tf32(TILE, 128);
            for (int tile = t0; tile < t1 && !s.abort_flag; ++tile) {
                const int it = tile - t0, buf = it & 1;
                if (!umma::mbar_wait(&s.acc_empty[buf], ((it >> 1) & 1) ^ 1)) s.abort_flag = 1;
                umma::fence_after_sync();
                const uint32_t tT = tmem + buf * 256, tH = tT + (WIDE ? 128 : 64);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!umma::mbar_wait(&s.full[h], it & 1)) s.abort_flag = 1;
                    umma::fence_after_sync();
                    const uint32_t aT = umma::smem_u32(s.A[h]), aH = aT + 2 * ARR_BYTES;
                    const uint32_t bW = umma::smem_u32(s.B[h]);
#pragma unroll
                    for (int ks = 0; ks < KH / 8; ++ks) {
                        const uint32_t off = ks * 32;                 // K = 8 fp32 inside the 128-byte row
                        const uint64_t dWhi = umma::make_smem_desc_sw128(bW + off);                // rows 0..63
                        const uint64_t dWlo = umma::make_smem_desc_sw128(bW + 64 * 128 + off);     // rows 64..127
                        const uint64_t dThi = umma::make_smem_desc_sw128(aT + off);
                        const uint64_t dHhi = umma::make_smem_desc_sw128(aH + off);
                        const uint32_t acc = (h | ks) != 0;
                        if constexpr (WIDE) {
                            umma::mma_tf32(tT, dThi, dWhi, idesc128, acc);
                            umma::mma_tf32(tT, umma::make_smem_desc_sw128(aT + ARR_BYTES + off), dWhi, idesc, 1);
                            umma::mma_tf32(tH, dHhi, dWhi, idesc128, acc);
                            umma::mma_tf32(tH, umma::make_smem_desc_sw128(aH + ARR_BYTES + off), dWhi, idesc, 1);
                        } else {
                            umma::mma_tf32(tT, dThi, dWhi, idesc, acc);
                            umma::mma_tf32(tT, dThi, dWlo, idesc, 1);
                            umma::mma_tf32(tT, umma::make_smem_desc_sw128(aT + ARR_BYTES + off), dWhi, idesc, 1);
                            umma::mma_tf32(tH, dHhi, dWhi, idesc, acc);
                            umma::mma_tf32(tH, dHhi, dWlo, idesc, 1);
                            umma::mma_tf32(tH, umma::make_smem_desc_sw128(aH + ARR_BYTES + off), dWhi, idesc, 1);
                        }
                    }
                    umma::mma_commit(&s.empty[h]);                   // stage h may be refilled once these finish
                }
                umma::mma_commit(&s.acc_full[buf]);                  // accumulator of this tile is complete
            }
        }
    } else {
        // =========================================================== epilogue (warps 8..11)
        const int q = warp & 3;                                       // TMEM lane quarter of this warp
        const int row = q * 32 + lane;
        const int et = tid - N_PROD;                                  // 0..127
        int rel = 0;
        for (int tile = t0; tile < t1; ++tile) {
            const int it = tile - t0, buf = it & 1;
            const TileInfo t = tile_info(s.prefix, rel_ptr, tile, rel);
            if (et < D) s.er[buf][et] = __ldg(rel_emb + static_cast<size_t>(t.rel) * D + et);
            const int pos = row < t.n_e ? __ldg(bk_pos + t.e0 + row) : -1;
            if (!umma::mbar_wait(&s.acc_full[buf], (it >> 1) & 1)) { s.abort_flag = 1; s.epi_abort = 1; }
            if (et == 0 && s.abort_flag) s.epi_abort = 1;
            umma::fence_after_sync();
            asm volatile("bar.sync 1, %0;" ::"n"(N_EPI) : "memory");  // er[buf] visible, abort view agreed
            if (s.epi_abort) break;
            const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
            float p = 0.f;
            if constexpr (WIDE) {
#pragma unroll
                for (int c0 = 0; c0 < D; c0 += 16) {
                    float a[16], b[16], c[16], d[16];
                    umma::tmem_ld16(ta + c0, a);                 // T: hi*hi + lo*hi
                    umma::tmem_ld16(ta + 64 + c0, b);            // T: hi*lo
                    umma::tmem_ld16(ta + 128 + c0, c);           // H: hi*hi + lo*hi
                    umma::tmem_ld16(ta + 192 + c0, d);           // H: hi*lo
                    umma::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        p = fmaf(a[i] + b[i], tanh_fast(c[i] + d[i] + s.er[buf][c0 + i]), p);
                }
            } else
#pragma unroll
            for (int c0 = 0; c0 < D; c0 += EPI_W) {
                float a[EPI_W], c[EPI_W];
                if constexpr (EPI_W == 32) {
                    umma::tmem_ld32(ta + c0, a);             // e_t W_r
                    umma::tmem_ld32(ta + 64 + c0, c);        // e_h W_r
                } else {
                    umma::tmem_ld16(ta + c0, a);
                    umma::tmem_ld16(ta + 64 + c0, c);
                }
                umma::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < EPI_W; ++i) p = fmaf(a[i], tanh_fast(c[i] + s.er[buf][c0 + i]), p);
            }
            umma::fence_before_sync();
            umma::mbar_arrive(&s.acc_empty[buf]);                     // TMEM buffer may be overwritten
            if (pos >= 0) logits_csr[pos] = p;
        }
    }

    umma::fence_before_sync();
    __syncthreads();
    if (tid == 0 && s.abort_flag) atomicExch(err, 1);
    if (warp == MMA_WARP) umma::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

__device__ int g_umma_error = 0;      // set when a tensor-core completion barrier timed out
// [0] = "the split-fp16 attention kernel cannot serve this call" (rewritten by every call), [1] = number of
// calls that therefore ran the fp32-range 3xTF32 kernel since the last kgat_attention_fallbacks()
__device__ int g_att_fallback[2] = {0, 0};

namespace {
// word[0] |= any |W_r| >= 32 (2^11 w_hi would overflow fp16, attention_umma16.cu); the launcher zeroes the word
// first.  A grid of CTAs: one CTA walked the 670 KB of an Amazon-book W_R in 10 us.
__global__ void __launch_bounds__(256) att_range_kernel(const float* __restrict__ W, int64_t n, int* __restrict__ word) {
    int bad = 0;
    const int64_t n4 = n >> 2, stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    for (int64_t i = t; i < n4; i += stride) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(W) + i);
        bad |= !(fabsf(v.x) < 32.0f) | !(fabsf(v.y) < 32.0f) | !(fabsf(v.z) < 32.0f) | !(fabsf(v.w) < 32.0f);
    }
    for (int64_t i = (n4 << 2) + t; i < n; i += stride) bad |= !(fabsf(__ldg(W + i)) < 32.0f);
    bad = __syncthreads_or(bad);
    if (threadIdx.x == 0 && bad) atomicOr(word, 1);
}
}  // namespace

}  // namespace kgat

using namespace kgat;

int kgat_eval_umma_error_take();

// device address of the shared tensor-core error word (attention_umma16.cu reports through it too)
int* kgat_umma_error_word() {
    static PerDevice<int*> cache;
    int*& p = cache.get();
    if (!p && cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_umma_error) != cudaSuccess) return nullptr;
    return p;
}

int* kgat_att_fallback_word() {
    static PerDevice<int*> cache;
    int*& p = cache.get();
    if (!p && cudaGetSymbolAddress(reinterpret_cast<void**>(&p), g_att_fallback) != cudaSuccess) return nullptr;
    return p;
}

int kgat_launch_att_range(const float* W_R, int64_t n, int* word, cudaStream_t st) {
    KGAT_CUDA(cudaMemsetAsync(word, 0, sizeof(int), st));
    const int64_t want = (n / 4 + 4 * 256 - 1) / (4 * 256);         // about four 16-byte loads per thread
    att_range_kernel<<<static_cast<unsigned>(want < 1 ? 1 : (want > 256 ? 256 : want)), 256, 0, st>>>(W_R, n, word);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_attention_fallbacks(void) {
    int v[2] = {0, 0}, zero = 0;
    if (cudaMemcpyFromSymbol(v, g_att_fallback, sizeof(v)) != cudaSuccess) return KGAT_ERR_CUDA;
    if (v[1] != 0 && cudaMemcpyToSymbol(g_att_fallback, &zero, sizeof(int), sizeof(int)) != cudaSuccess)
        return KGAT_ERR_CUDA;
    return v[1];
}

extern "C" int kgat_umma_error_flag(void) {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_umma_error, sizeof(int)) != cudaSuccess) return KGAT_ERR_CUDA;
    if (v != 0) cudaMemcpyToSymbol(g_umma_error, &zero, sizeof(int));
    const int e = kgat_eval_umma_error_take();
    if (e < 0) return e;
    return (v != 0 || e != 0) ? 1 : 0;
}

// Launcher used by kgat_attention_logits for d = rd = 64 (declared in attention.cu).
int kgat_launch_att_logits_umma(const kgat_graph_t* g, const float* entity, const float* rel_emb, const float* W_R,
                                float* logits_csr, int* only_if, cudaStream_t st) {
    static PerDevice<int*> err_flags;
    static PerDeviceOnce configured;
    static int epi_w = 32;
    if (configured.need()) {
        KGAT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&err_flags.get()), g_umma_error));
        KGAT_CUDA(cudaFuncSetAttribute(att_logits_umma_kernel<32, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
        KGAT_CUDA(cudaFuncSetAttribute(att_logits_umma_kernel<16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
        const char* e = std::getenv("KGAT_ATT_EPI");
        if (e && std::atoi(e) == 16) epi_w = 16;
        configured.set();
    }
    const int64_t max_tiles = g->n_edges / TILE + g->n_rel;
    const int grid = static_cast<int>(max_tiles < num_sms() ? max_tiles : num_sms());
    int* err_flag = err_flags.get();
    auto kern = epi_w == 16 ? att_logits_umma_kernel<16, 1> : att_logits_umma_kernel<32, 0>;
    kern<<<grid, THREADS, sizeof(Smem), st>>>(g->rel_ptr, static_cast<int>(g->n_rel), g->bk_head, g->bk_tail,
                                              g->bk_pos, entity, rel_emb, W_R, logits_csr, err_flag, only_if);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
