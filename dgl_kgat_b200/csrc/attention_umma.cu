// (2) TransR attention logits on the 5th-gen tensor cores: relation-grouped projection as a
// tcgen05 GEMM in 3xTF32, fused with the per-edge gather (prologue) and the tanh / dot epilogue.
//
//   pi(h,r,t) = (e_t W_r) . tanh(e_h W_r + e_r)                        (reference models.py:135-144)
//
// One persistent CTA per SM walks a contiguous range of 128-edge tiles of the relation buckets.
// Per tile:
//   gather   the 128 tail rows and 128 head rows of the entity table (128-bit loads, 8 rows x 64 B per
//            warp instruction), split every value x = hi + lo (hi = RN-TF32(x), lo exact) and store
//            both halves in shared memory in the K-major no-swizzle core-matrix layout (umma.cuh);
//   MMA      A = gathered rows (M = 128 edges, K = 64), B = [W_hi | W_lo] (N = 128, staged once per
//            relation): T_hi x [W_hi|W_lo] (N=128) and T_lo x W_hi (N=64) accumulate hi*hi + lo*hi in
//            TMEM columns 0..63 and hi*lo in columns 64..127; same for the head rows.  fp32
//            accumulation in TMEM; 32 tcgen05.mma per tile issued by one thread;
//   epilogue (overlaps the next tile's MMAs, TMEM double-buffered) tcgen05.ld, t = c[0:64]+c[64:128],
//            logit = sum_c t_r[c] * tanh(h_r[c] + e_r[c]), scattered to the edge's CSR position.
// Neither projection ever exists in HBM.  3xTF32 keeps the result inside the 1e-5 relative bar
// (a single TF32 pass is ~5e-4).
#include "common.cuh"
#include "umma.cuh"

namespace kgat {
namespace {

constexpr int TILE = 128;
constexpr int D = 64;            // entity and relation width of this specialisation
constexpr int THREADS = 256;
constexpr int MAX_REL = 1024;
constexpr uint32_t A_BYTES = TILE * D * 4;          // 32 KB per operand array
constexpr uint32_t SBO = (D / 4) * 128;             // 8-row group stride for kt = 64
constexpr uint32_t TMEM_COLS = 512;                 // 2 buffers x (T: 128 + H: 128)

struct Smem {
    unsigned char B[2 * 64 * D * 4];                // rows 0..63: W_hi^T, rows 64..127: W_lo^T
    unsigned char A[4][A_BYTES];                    // T_hi, T_lo, H_hi, H_lo
    float er[2][D];
    float part[TILE];
    int tail[TILE], head[TILE];
    int pos[2][TILE];
    int prefix[MAX_REL + 1];
    uint64_t mma_bar[2];
    uint32_t tmem_base;
    int cur_rel_loaded;
};

// tanh to ~3e-7 absolute: odd polynomial below 0.1, 1 - 2 / (1 + exp(2|x|)) above
__device__ __forceinline__ float tanh_fast(float x) {
    const float ax = fabsf(x);
    const float x2 = x * x;
    const float poly = x * fmaf(x2, fmaf(x2, fmaf(x2, -0.05396825f, 0.13333334f), -0.33333334f), 1.0f);
    const float e = __expf(2.0f * ax);
    const float big = copysignf(1.0f - __fdividef(2.0f, 1.0f + e), x);
    return ax < 0.1f ? poly : big;
}

__global__ void __launch_bounds__(THREADS, 1)
att_logits_umma_kernel(const int32_t* __restrict__ rel_ptr, int n_rel, const int32_t* __restrict__ bk_head,
                       const int32_t* __restrict__ bk_tail, const int32_t* __restrict__ bk_pos,
                       const float* __restrict__ ent, const float* __restrict__ rel_emb,
                       const float* __restrict__ W_R, float* __restrict__ logits_csr,
                       int* __restrict__ err) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& s = *reinterpret_cast<Smem*>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

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
    if (warp == 1) umma::tmem_alloc(&s.tmem_base, TMEM_COLS);
    if (tid == 64) {
        umma::mbar_init(&s.mma_bar[0], 1);
        umma::mbar_init(&s.mma_bar[1], 1);
        umma::fence_barrier_init();
        s.cur_rel_loaded = -1;
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = s.tmem_base;
    const int n_tiles = s.prefix[n_rel];
    const int t0 = static_cast<int>(static_cast<long long>(n_tiles) * blockIdx.x / gridDim.x);
    const int t1 = static_cast<int>(static_cast<long long>(n_tiles) * (blockIdx.x + 1) / gridDim.x);

    // gather mapping: a warp instruction covers 8 rows x 4 consecutive 16-byte chunks
    const int r8 = lane & 7, cq = lane >> 3;
    const uint32_t lbo = umma::LBO_BYTES, sbo = SBO;
    const uint32_t idesc128 = umma::make_idesc_tf32(TILE, 128), idesc64 = umma::make_idesc_tf32(TILE, 64);

    int rel = 0;
    int prev_ne = 0;
    bool failed = false;

    // epilogue of the tile whose accumulators sit in TMEM buffer `buf`
    auto epilogue = [&](int buf, int n_e) {
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane, c0 = half * 32;
        const uint32_t ta = tmem + (static_cast<uint32_t>(q * 32) << 16) + buf * 256;
        float a[32], b[32], t[32];
        umma::tmem_ld32(ta + c0, a);
        umma::tmem_ld32(ta + 64 + c0, b);
#pragma unroll
        for (int i = 0; i < 32; ++i) t[i] = a[i] + b[i];
        umma::tmem_ld32(ta + 128 + c0, a);
        umma::tmem_ld32(ta + 192 + c0, b);
        float p = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) p = fmaf(t[i], tanh_fast(a[i] + b[i] + s.er[buf][c0 + i]), p);
        if (half == 1) s.part[row] = p;
        umma::fence_before_sync();
        __syncthreads();
        if (half == 0 && row < n_e) logits_csr[s.pos[buf][row]] = p + s.part[row];
    };

    for (int tile = t0; tile < t1; ++tile) {
        const int it = tile - t0, buf = it & 1;
        while (tile >= s.prefix[rel + 1]) ++rel;                       // uniform across the CTA
        const int e0 = __ldg(rel_ptr + rel) + (tile - s.prefix[rel]) * TILE;
        const int n_e = min(TILE, __ldg(rel_ptr + rel + 1) - e0);

        // (a) edge ids of this tile
        if (tid < TILE) {
            const bool ok = tid < n_e;
            s.tail[tid] = ok ? __ldg(bk_tail + e0 + tid) : -1;
            s.head[tid] = ok ? __ldg(bk_head + e0 + tid) : -1;
            s.pos[buf][tid] = ok ? __ldg(bk_pos + e0 + tid) : -1;
        }
        __syncthreads();

        // (b) issue the row gather into registers: 16 x 128-bit loads in flight per thread
        float4 v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int combo = warp * 16 + j;
            const int rg = combo >> 2, cg = combo & 3;                 // row group 0..31 (16.. = head rows)
            const int rl = (rg & 15) * 8 + r8;
            const int id = rg < 16 ? s.tail[rl] : s.head[rl];
            v[j] = id >= 0 ? ldg4(ent + static_cast<size_t>(id) * D + (cg * 4 + cq) * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
        }

        // (c) the previous tile's MMAs must have finished reading shared memory
        if (it > 0) {
            const bool ok = umma::mbar_wait(&s.mma_bar[buf ^ 1], ((it - 1) >> 1) & 1);
            failed = __syncthreads_or(!ok) != 0;          // CTA-uniform: a timeout aborts the whole CTA
            umma::fence_after_sync();
            if (failed) break;
        }

        // (d) restage W_r / e_r when the relation changes
        if (s.cur_rel_loaded != rel) {
            const float* Wg = W_R + static_cast<size_t>(rel) * D * D;
            for (int idx = tid; idx < D * D; idx += THREADS) {
                const int k = idx >> 6, n = idx & 63;
                float hi, lo;
                umma::split_tf32(__ldg(Wg + idx), hi, lo);
                *reinterpret_cast<float*>(s.B + umma::tile_offset(n, k, D)) = hi;
                *reinterpret_cast<float*>(s.B + umma::tile_offset(64 + n, k, D)) = lo;
            }
        }
        if (tid < D) s.er[buf][tid] = __ldg(rel_emb + static_cast<size_t>(rel) * D + tid);

        // (e) split and store the gathered rows
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int combo = warp * 16 + j;
            const int rg = combo >> 2, cg = combo & 3;
            const int rl = (rg & 15) * 8 + r8;
            const uint32_t off = (rl >> 3) * SBO + (cg * 4 + cq) * 128 + r8 * 16;
            float4 hi, lo;
            umma::split_tf32(v[j].x, hi.x, lo.x);
            umma::split_tf32(v[j].y, hi.y, lo.y);
            umma::split_tf32(v[j].z, hi.z, lo.z);
            umma::split_tf32(v[j].w, hi.w, lo.w);
            unsigned char* base = rg < 16 ? s.A[0] : s.A[2];
            *reinterpret_cast<float4*>(base + off) = hi;
            *reinterpret_cast<float4*>(base + A_BYTES + off) = lo;
        }
        umma::fence_proxy_async();
        umma::fence_before_sync();
        __syncthreads();
        umma::fence_after_sync();
        if (tid == 0) s.cur_rel_loaded = rel;

        // (f) one thread issues the tile's 32 MMAs
        if (tid == 0) {
            const uint32_t tT = tmem + buf * 256, tH = tT + 128;
            const uint32_t aT = umma::smem_u32(s.A[0]), aH = umma::smem_u32(s.A[2]), bW = umma::smem_u32(s.B);
#pragma unroll
            for (int ks = 0; ks < D / 8; ++ks) {
                const uint32_t off = ks * 2 * umma::CORE_BYTES;
                const uint64_t dB = umma::make_smem_desc(bW + off, lbo, sbo);
                umma::mma_tf32(tT, umma::make_smem_desc(aT + off, lbo, sbo), dB, idesc128, ks > 0);
                umma::mma_tf32(tT, umma::make_smem_desc(aT + A_BYTES + off, lbo, sbo), dB, idesc64, 1);
                umma::mma_tf32(tH, umma::make_smem_desc(aH + off, lbo, sbo), dB, idesc128, ks > 0);
                umma::mma_tf32(tH, umma::make_smem_desc(aH + A_BYTES + off, lbo, sbo), dB, idesc64, 1);
            }
            umma::mma_commit(&s.mma_bar[buf]);
        }

        // (g) epilogue of the previous tile overlaps these MMAs
        if (it > 0) epilogue(buf ^ 1, prev_ne);
        prev_ne = n_e;
    }

    // drain: last tile
    if (t1 > t0 && !failed) {
        const int it = t1 - t0 - 1, buf = it & 1;
        const bool ok = umma::mbar_wait(&s.mma_bar[buf], (it >> 1) & 1);
        failed = __syncthreads_or(!ok) != 0;
        umma::fence_after_sync();
        if (!failed) epilogue(buf, prev_ne);
    }
    if (failed && tid == 0) atomicExch(err, 1);
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace
}  // namespace kgat

namespace kgat {
__device__ int g_umma_error = 0;      // set when a tensor-core completion barrier timed out
}

using namespace kgat;

extern "C" int kgat_umma_error_flag(void) {
    int v = 0, zero = 0;
    if (cudaMemcpyFromSymbol(&v, g_umma_error, sizeof(int)) != cudaSuccess) return KGAT_ERR_CUDA;
    if (v != 0) cudaMemcpyToSymbol(g_umma_error, &zero, sizeof(int));
    return v;
}

// Launcher used by kgat_attention_logits for d = rd = 64 (declared in attention.cu).
int kgat_launch_att_logits_umma(const kgat_graph_t* g, const float* entity, const float* rel_emb, const float* W_R,
                                float* logits_csr, cudaStream_t st) {
    int* err_flag = nullptr;
    KGAT_CUDA(cudaGetSymbolAddress(reinterpret_cast<void**>(&err_flag), g_umma_error));
    static bool configured = false;
    if (!configured) {
        KGAT_CUDA(cudaFuncSetAttribute(att_logits_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(Smem)));
        configured = true;
    }
    const int64_t max_tiles = g->n_edges / TILE + g->n_rel;
    const int grid = static_cast<int>(max_tiles < num_sms() ? max_tiles : num_sms());
    att_logits_umma_kernel<<<grid, THREADS, sizeof(Smem), st>>>(
        g->rel_ptr, static_cast<int>(g->n_rel), g->bk_head, g->bk_tail, g->bk_pos, entity, rel_emb, W_R, logits_csr,
        err_flag);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
