// (6) TransR step of the KG phase: loss AND every parameter gradient in one pass.
//
// Replaces Model.transR (reference models.py:114-133) + its autograd:
//   v_x = normalize(e_x W_r)  for x in {h, pos_t, neg_t},   r_v = normalize(e_r)
//   loss = mean(-logsigmoid(|v_h + r_v - v_n|^2 - |v_h + r_v - v_p|^2))
//        + lambda * mean over the batch of (|v_h|^2 + |r_v|^2 + |v_p|^2 + |v_n|^2) / 2
// The reference gathers W_R[r] per sample (a (B, d, rd) tensor, models.py:46-47) and runs three bmm;
// here the batch is grouped by relation so W_r is staged once per 32-sample chunk:
//   plan     (one CTA) stable partition of the samples by relation (per-warp match/ballot counts) ->
//            chunk table; bitonic sort of (node, gradient row) for the 3B entity rows -> the order
//            their gradients are summed in;
//   main     one CTA per chunk: gather the 96 entity rows, V = E W_r, per-sample normalise / scores /
//            loss / gradient of the projections, dE = G W_r^T (rows written per sample),
//            dW_r partial = E^T G, d e_r partial, loss partial; the last CTA to finish adds the loss
//            partials in chunk order;
//   reduce   (backward) per relation, partials added in chunk order (fp64), times the upstream
//            gradient -> d_W, d_rel;
//   scatter  (backward) one warp per distinct node adds its gradient rows in sorted order -> d_entity.
// Deterministic: every sum has a fixed order (the one atomic is the CTA-completion counter).
// d = rd = 64, 3 * batch <= 16384.
#include "common.cuh"

namespace kgat {
namespace {

constexpr int D = 64;
constexpr int T = 32;                   // samples per chunk
constexpr int THREADS = 256;
constexpr int LDW = D + 1;              // padded W_r rows: conflict-free along either index
constexpr int MAX_REL = 1024;
constexpr int MAX_ROWS = 16384;         // 3 * batch, bitonic-sorted in shared memory (8 B keys)

struct Ws {
    int32_t *order_rel, *chunk_rel, *chunk_start, *chunk_cnt, *rel_chunk0, *n_chunks, *ent_node, *ent_row;
    unsigned* done;
    float *dE_rows, *dW_part, *drel_part;
    double* loss_part;
    size_t bytes;
};

inline size_t align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }

Ws carve(void* base, int64_t B, int64_t R) {
    const int64_t maxc = B / T + R + 1, P = 3 * B;
    unsigned char* p = static_cast<unsigned char*>(base);
    size_t off = 0;
    auto take = [&](size_t n) { unsigned char* q = p ? p + off : nullptr; off += align16(n); return q; };
    Ws w;
    w.order_rel = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * B));
    w.chunk_rel = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * maxc));
    w.chunk_start = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * maxc));
    w.chunk_cnt = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * maxc));
    w.rel_chunk0 = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * (R + 1)));
    w.n_chunks = reinterpret_cast<int32_t*>(take(sizeof(int32_t)));
    w.done = reinterpret_cast<unsigned*>(take(sizeof(unsigned)));
    w.ent_node = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * P));
    w.ent_row = reinterpret_cast<int32_t*>(take(sizeof(int32_t) * P));
    w.dE_rows = reinterpret_cast<float*>(take(sizeof(float) * P * D));
    w.dW_part = reinterpret_cast<float*>(take(sizeof(float) * maxc * D * D));
    w.drel_part = reinterpret_cast<float*>(take(sizeof(float) * maxc * D));
    w.loss_part = reinterpret_cast<double*>(take(sizeof(double) * maxc));
    w.bytes = off;
    return w;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
constexpr int PLAN_THREADS = 1024;
constexpr int PLAN_WARPS = PLAN_THREADS / 32;
constexpr int MAX_ROUNDS = (MAX_ROWS / 3 + PLAN_THREADS - 1) / PLAN_THREADS;     // samples per plan thread

// Bitonic sort of n (power of two, >= 32 * ... handled below) DISTINCT keys in shared memory by the whole
// CTA.  Compare-exchange distances below 32 stay inside a warp: those steps run on registers with
// shuffles (element e lives in thread e % 1024, slot e / 1024), so only the long distances pay a
// __syncthreads; the shared-memory steps enumerate PAIRS, so no thread idles.
template <typename K>
__device__ __forceinline__ K shfl_xor_key(K v, int j) {
    if constexpr (sizeof(K) == 8) {
        const unsigned lo = __shfl_xor_sync(FULL_MASK, static_cast<unsigned>(v), j);
        const unsigned hi = __shfl_xor_sync(FULL_MASK, static_cast<unsigned>(v >> 32), j);
        return (static_cast<K>(hi) << 32) | lo;
    } else {
        return __shfl_xor_sync(FULL_MASK, v, j);
    }
}

template <typename K>
__device__ void bitonic_sort(K* keys, int n) {
    const int tid = threadIdx.x;
    for (int k = 2; k <= n; k <<= 1) {
        int j = k >> 1;
        for (; j >= 32; j >>= 1) {                       // long distances: shared memory, one pair per step
            for (int t = tid; t < (n >> 1); t += PLAN_THREADS) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), x = i | j;
                const K a = keys[i], b = keys[x];
                if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[x] = a; }
            }
            __syncthreads();
        }
        // distances 16..1 (or k/2..1 for the first stages): registers + shuffles
        for (int e = tid; e < n; e += PLAN_THREADS) {
            K v = keys[e];
            const bool up = (e & k) == 0;
            for (int jj = j; jj > 0; jj >>= 1) {
                const K o = shfl_xor_key<K>(v, jj);
                const bool lower = (e & jj) == 0;
                v = (lower == up) ? (v < o ? v : o) : (v < o ? o : v);
            }
            keys[e] = v;
        }
        __syncthreads();
    }
}

// Shared memory is used twice: first as the per-warp relation counters of the stable partition, then as
// the key array of the entity-row sort.
template <typename K>
__global__ void __launch_bounds__(PLAN_THREADS)
transr_plan_kernel(const int64_t* __restrict__ h, const int64_t* __restrict__ r, const int64_t* __restrict__ pos,
                   const int64_t* __restrict__ neg, int B, int R, int n2_ent, int idx_bits,
                   int32_t* __restrict__ order_rel, int32_t* __restrict__ chunk_rel,
                   int32_t* __restrict__ chunk_start, int32_t* __restrict__ chunk_cnt,
                   int32_t* __restrict__ rel_chunk0, int32_t* __restrict__ n_chunks, unsigned* __restrict__ done,
                   int32_t* __restrict__ ent_node, int32_t* __restrict__ ent_row) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int sStart[MAX_REL + 1], sC0[MAX_REL + 1], sTot[MAX_REL];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) *done = 0;

    // ---- samples grouped by relation, original order kept inside a relation (stable partition):
    // warp w owns the contiguous samples [w * per, (w + 1) * per); cnt[w][rel] counts what it has seen.
    int* cnt = reinterpret_cast<int*>(smem_raw);
    for (int i = tid; i < PLAN_WARPS * R; i += PLAN_THREADS) cnt[i] = 0;
    __syncthreads();
    const int per = ((B + PLAN_WARPS - 1) / PLAN_WARPS + 31) & ~31;
    const int wbeg = warp * per, wend = min(B, wbeg + per);
    int my_rel[MAX_ROUNDS], my_pref[MAX_ROUNDS];
#pragma unroll
    for (int rd = 0; rd < MAX_ROUNDS; ++rd) {
        const int i = wbeg + rd * 32 + lane;
        const bool valid = rd * 32 < per && i < wend;
        const int rel = valid ? static_cast<int>(r[i]) : -1;
        const unsigned m = __match_any_sync(FULL_MASK, rel);
        const int rank = __popc(m & ((1u << lane) - 1u));
        int prev = 0;
        if (valid) prev = cnt[warp * R + rel];
        __syncwarp();
        if (valid && rank == 0) cnt[warp * R + rel] = prev + __popc(m);
        __syncwarp();
        my_rel[rd] = rel;
        my_pref[rd] = prev + rank;
    }
    __syncthreads();
    for (int rr = tid; rr < R; rr += PLAN_THREADS) {     // exclusive scan over the warps, per relation
        int run = 0;
        for (int w = 0; w < PLAN_WARPS; ++w) {
            const int t = cnt[w * R + rr];
            cnt[w * R + rr] = run;
            run += t;
        }
        sTot[rr] = run;
    }
    __syncthreads();
    if (warp == 0) {                                     // scan over relations: lane l takes a contiguous block
        const int blk = (R + 31) / 32, b0 = min(R, lane * blk), b1 = min(R, b0 + blk);
        int tot = 0, totc = 0;
        for (int rr = b0; rr < b1; ++rr) { tot += sTot[rr]; totc += (sTot[rr] + T - 1) / T; }
        int inc = tot, incc = totc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL_MASK, inc, o), vc = __shfl_up_sync(FULL_MASK, incc, o);
            if (lane >= o) { inc += v; incc += vc; }
        }
        int run = inc - tot, runc = incc - totc;
        for (int rr = b0; rr < b1; ++rr) {
            sStart[rr] = run;
            sC0[rr] = runc;
            run += sTot[rr];
            runc += (sTot[rr] + T - 1) / T;
        }
        if (lane == 31) { sStart[R] = inc; sC0[R] = incc; *n_chunks = incc; }
    }
    __syncthreads();
#pragma unroll
    for (int rd = 0; rd < MAX_ROUNDS; ++rd) {
        const int i = wbeg + rd * 32 + lane;
        if (my_rel[rd] >= 0) order_rel[sStart[my_rel[rd]] + cnt[warp * R + my_rel[rd]] + my_pref[rd]] = i;
    }
    for (int rr = tid; rr <= R; rr += PLAN_THREADS) rel_chunk0[rr] = sC0[rr];
    for (int rr = tid; rr < R; rr += PLAN_THREADS) {
        const int s = sStart[rr], e = sStart[rr + 1];
        for (int c = sC0[rr], b = s; b < e; ++c, b += T) {
            chunk_rel[c] = rr;
            chunk_start[c] = b;
            chunk_cnt[c] = min(T, e - b);
        }
    }
    __syncthreads();

    // ---- entity gradient rows by (node, row): row = set * B + sample
    K* keys = reinterpret_cast<K*>(smem_raw);
    const int P = 3 * B;
    for (int i = tid; i < n2_ent; i += PLAN_THREADS) {
        K k = ~static_cast<K>(0);
        if (i < P) {
            const int64_t node = i < B ? h[i] : (i < 2 * B ? pos[i - B] : neg[i - 2 * B]);
            k = (static_cast<K>(node) << idx_bits) | static_cast<K>(i);
        }
        keys[i] = k;
    }
    __syncthreads();
    bitonic_sort<K>(keys, n2_ent);
    const K idx_mask = (static_cast<K>(1) << idx_bits) - 1;
    for (int i = tid; i < P; i += PLAN_THREADS) {
        ent_node[i] = static_cast<int32_t>(keys[i] >> idx_bits);
        ent_row[i] = static_cast<int32_t>(keys[i] & idx_mask);
    }
}

// ------------------------------------------------------------------------------------------------
// main
// ------------------------------------------------------------------------------------------------
struct MainSmem {
    float W[D * LDW];
    float E[3 * T * D];
    float G[3 * T * D];       // projections, then their gradients
    float rv[D];
    float dr[THREADS / 32][D];
    double loss[THREADS / 32];
    float rnorm;
    int idx[T];
};

// gradient through v = y / max(|y|, eps) for this lane's two elements
__device__ __forceinline__ float2 normalize_bwd(float2 v, float2 dv, float norm) {
    if (norm >= NORM_EPS) {
        const float dot = warp_sum(v.x * dv.x + v.y * dv.y);
        return make_float2((dv.x - v.x * dot) / norm, (dv.y - v.y * dot) / norm);
    }
    return make_float2(dv.x / NORM_EPS, dv.y / NORM_EPS);
}

__global__ void __launch_bounds__(THREADS)
transr_main_kernel(const float* __restrict__ entity, const float* __restrict__ rel_emb,
                   const float* __restrict__ W_R, const int64_t* __restrict__ h, const int64_t* __restrict__ pos,
                   const int64_t* __restrict__ neg, int B, float lambda,
                   const int32_t* __restrict__ order_rel, const int32_t* __restrict__ chunk_rel,
                   const int32_t* __restrict__ chunk_start, const int32_t* __restrict__ chunk_cnt,
                   const int32_t* __restrict__ n_chunks, unsigned* __restrict__ done, float* __restrict__ dE_rows,
                   float* __restrict__ dW_part, float* __restrict__ drel_part, double* __restrict__ loss_part,
                   float* __restrict__ loss) {
    const int c = blockIdx.x;
    const int nc = *n_chunks;
    if (c >= nc) return;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MainSmem& s = *reinterpret_cast<MainSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rel = chunk_rel[c], s0 = chunk_start[c], cnt = chunk_cnt[c];
    const float invB = 1.0f / static_cast<float>(B);

    // ---- a. stage W_r, the sample ids, the 96 entity rows, the relation vector
    const float* Wg = W_R + static_cast<size_t>(rel) * D * D;
    for (int i = tid; i < D * D; i += THREADS) s.W[(i >> 6) * LDW + (i & 63)] = __ldg(Wg + i);
    if (tid < T) s.idx[tid] = tid < cnt ? order_rel[s0 + tid] : -1;
    __syncthreads();
    for (int q = tid; q < 3 * T * (D / 4); q += THREADS) {
        const int row = q >> 4, c4 = q & 15, set = row / T, i = s.idx[row % T];
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i >= 0) {
            const int64_t node = set == 0 ? h[i] : (set == 1 ? pos[i] : neg[i]);
            v = ldg4(entity + static_cast<size_t>(node) * D + c4 * 4);
        }
        *reinterpret_cast<float4*>(s.E + row * D + c4 * 4) = v;
    }
    if (warp == 0) {
        const float2 er = *reinterpret_cast<const float2*>(rel_emb + static_cast<size_t>(rel) * D + lane * 2);
        const float n = sqrtf(warp_sum(er.x * er.x + er.y * er.y));
        const float dn = fmaxf(n, NORM_EPS);
        s.rv[lane * 2] = er.x / dn;
        s.rv[lane * 2 + 1] = er.y / dn;
        if (lane == 0) s.rnorm = n;
    }
    __syncthreads();

    // ---- b. V = E W_r  (96 x 64): thread = output column n, rows rg + 4 m
    {
        const int n = tid & 63, rg = tid >> 6;
        float acc[24];
#pragma unroll
        for (int m = 0; m < 24; ++m) acc[m] = 0.f;
#pragma unroll 2
        for (int k0 = 0; k0 < D; k0 += 4) {
            const float w0 = s.W[(k0 + 0) * LDW + n], w1 = s.W[(k0 + 1) * LDW + n], w2 = s.W[(k0 + 2) * LDW + n],
                        w3 = s.W[(k0 + 3) * LDW + n];
#pragma unroll
            for (int m = 0; m < 24; ++m) {
                const float4 e = *reinterpret_cast<const float4*>(s.E + (rg + 4 * m) * D + k0);
                acc[m] = fmaf(e.x, w0, acc[m]);
                acc[m] = fmaf(e.y, w1, acc[m]);
                acc[m] = fmaf(e.z, w2, acc[m]);
                acc[m] = fmaf(e.w, w3, acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < 24; ++m) s.G[(rg + 4 * m) * D + n] = acc[m];
    }
    __syncthreads();

    // ---- c. per sample: normalise, scores, loss, gradient of the three projections (in place in G)
    {
        const float2 rv = make_float2(s.rv[lane * 2], s.rv[lane * 2 + 1]);
        float2 dr_acc = make_float2(0.f, 0.f);
        double loss_acc = 0.0;
        for (int sl = warp; sl < T; sl += THREADS / 32) {
            float2* gh = reinterpret_cast<float2*>(s.G + (0 * T + sl) * D) + lane;
            float2* gp = reinterpret_cast<float2*>(s.G + (1 * T + sl) * D) + lane;
            float2* gn = reinterpret_cast<float2*>(s.G + (2 * T + sl) * D) + lane;
            if (sl >= cnt) {                                   // padding rows carry no gradient
                *gh = *gp = *gn = make_float2(0.f, 0.f);
                continue;
            }
            const float2 yh = *gh, yp = *gp, yn = *gn;
            const float nh = sqrtf(warp_sum(yh.x * yh.x + yh.y * yh.y));
            const float np = sqrtf(warp_sum(yp.x * yp.x + yp.y * yp.y));
            const float nn = sqrtf(warp_sum(yn.x * yn.x + yn.y * yn.y));
            const float dh = fmaxf(nh, NORM_EPS), dp = fmaxf(np, NORM_EPS), dn = fmaxf(nn, NORM_EPS);
            const float2 hv = make_float2(yh.x / dh, yh.y / dh), pv = make_float2(yp.x / dp, yp.y / dp),
                         nv = make_float2(yn.x / dn, yn.y / dn);
            const float2 u = make_float2(hv.x + rv.x - pv.x, hv.y + rv.y - pv.y);
            const float2 w = make_float2(hv.x + rv.x - nv.x, hv.y + rv.y - nv.y);
            const float ps = warp_sum(u.x * u.x + u.y * u.y), ns = warp_sum(w.x * w.x + w.y * w.y);
            const float x = ns - ps;
            const float reg = 0.5f * warp_sum(hv.x * hv.x + hv.y * hv.y + rv.x * rv.x + rv.y * rv.y +
                                              pv.x * pv.x + pv.y * pv.y + nv.x * nv.x + nv.y * nv.y);
            // -logsigmoid(x) = softplus(-x); its derivative w.r.t. x is -sigmoid(-x)
            const float l = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
            const float sg = 1.0f / (1.0f + expf(x));
            loss_acc += static_cast<double>(l) + static_cast<double>(lambda) * static_cast<double>(reg);
            const float a = 2.0f * sg * invB, lb = lambda * invB;
            // d loss / d pos_score = +sg / B, d loss / d neg_score = -sg / B
            const float2 du = make_float2(a * u.x, a * u.y), dw = make_float2(-a * w.x, -a * w.y);
            const float2 dhv = make_float2(du.x + dw.x + lb * hv.x, du.y + dw.y + lb * hv.y);
            const float2 drv = make_float2(du.x + dw.x + lb * rv.x, du.y + dw.y + lb * rv.y);
            const float2 dpv = make_float2(-du.x + lb * pv.x, -du.y + lb * pv.y);
            const float2 dnv = make_float2(-dw.x + lb * nv.x, -dw.y + lb * nv.y);
            *gh = normalize_bwd(hv, dhv, nh);
            *gp = normalize_bwd(pv, dpv, np);
            *gn = normalize_bwd(nv, dnv, nn);
            dr_acc.x += drv.x;
            dr_acc.y += drv.y;
        }
        s.dr[warp][lane * 2] = dr_acc.x;
        s.dr[warp][lane * 2 + 1] = dr_acc.y;
        if (lane == 0) s.loss[warp] = loss_acc;
    }
    __syncthreads();
    if (warp == 0) {
        float2 t = make_float2(0.f, 0.f);
        double ls = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) {
            t.x += s.dr[w][lane * 2];
            t.y += s.dr[w][lane * 2 + 1];
            ls += s.loss[w];
        }
        const float2 rv = make_float2(s.rv[lane * 2], s.rv[lane * 2 + 1]);
        const float2 de = normalize_bwd(rv, t, s.rnorm);
        *reinterpret_cast<float2*>(drel_part + static_cast<size_t>(c) * D + lane * 2) = de;
        if (lane == 0) loss_part[c] = ls;
    }

    // ---- d. dE = G W_r^T, written per (set, sample) row
    {
        const int k = tid & 63, rg = tid >> 6;
        float acc[24];
#pragma unroll
        for (int m = 0; m < 24; ++m) acc[m] = 0.f;
#pragma unroll 2
        for (int n0 = 0; n0 < D; n0 += 4) {
            const float w0 = s.W[k * LDW + n0], w1 = s.W[k * LDW + n0 + 1], w2 = s.W[k * LDW + n0 + 2],
                        w3 = s.W[k * LDW + n0 + 3];
#pragma unroll
            for (int m = 0; m < 24; ++m) {
                const float4 g = *reinterpret_cast<const float4*>(s.G + (rg + 4 * m) * D + n0);
                acc[m] = fmaf(g.x, w0, acc[m]);
                acc[m] = fmaf(g.y, w1, acc[m]);
                acc[m] = fmaf(g.z, w2, acc[m]);
                acc[m] = fmaf(g.w, w3, acc[m]);
            }
        }
#pragma unroll
        for (int m = 0; m < 24; ++m) {
            const int row = rg + 4 * m, set = row / T, i = s.idx[row % T];
            if (i >= 0) dE_rows[(static_cast<size_t>(set) * B + i) * D + k] = acc[m];
        }
    }

    // ---- e. dW_r partial = E^T G over the 96 rows: thread = 4 x 4 block (k0.., n0..)
    {
        const int k0 = (tid & 15) * 4, n0 = (tid >> 4) * 4;
        float acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
#pragma unroll 4
        for (int row = 0; row < 3 * T; ++row) {
            const float4 e = *reinterpret_cast<const float4*>(s.E + row * D + k0);
            const float4 g = *reinterpret_cast<const float4*>(s.G + row * D + n0);
            const float ev[4] = {e.x, e.y, e.z, e.w}, gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(ev[a], gv[b], acc[a][b]);
        }
        float* out = dW_part + static_cast<size_t>(c) * D * D;
#pragma unroll
        for (int a = 0; a < 4; ++a)
            *reinterpret_cast<float4*>(out + (k0 + a) * D + n0) = make_float4(acc[a][0], acc[a][1], acc[a][2], acc[a][3]);
    }

    // ---- the last chunk to finish adds the loss partials, always in chunk order
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (tid == 0) is_last = atomicAdd(done, 1u) == static_cast<unsigned>(nc - 1);
    __syncthreads();
    if (is_last && warp == 0) {
        __threadfence();
        double a = 0.0;
        for (int c0 = 0; c0 < nc; c0 += 32) {            // 32 partials per pass, combined in a fixed tree
            double v = c0 + lane < nc ? __ldcg(loss_part + c0 + lane) : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
            a += v;
        }
        if (lane == 0) *loss = static_cast<float>(a / static_cast<double>(B));
    }
}

// ------------------------------------------------------------------------------------------------
// backward.  reduce: grid (R, 4); CTA (rel, quarter) sums its quarter of d_W[rel] over the relation's
// chunks.  grad: device scalar, the upstream gradient of the loss (null = 1).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
transr_reduce_kernel(const int32_t* __restrict__ rel_chunk0, const float* __restrict__ dW_part,
                     const float* __restrict__ drel_part, const float* __restrict__ grad,
                     float* __restrict__ d_W, float* __restrict__ d_rel) {
    const int rel = blockIdx.x, quarter = blockIdx.y, tid = threadIdx.x;
    const int c0 = rel_chunk0[rel], c1 = rel_chunk0[rel + 1];
    const double g = grad ? static_cast<double>(*grad) : 1.0;
    const int e = quarter * 1024 + tid * 4;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    for (int c = c0; c < c1; ++c) {
        const float4 v = ldg4(dW_part + static_cast<size_t>(c) * D * D + e);
        a0 += v.x; a1 += v.y; a2 += v.z; a3 += v.w;
    }
    *reinterpret_cast<float4*>(d_W + static_cast<size_t>(rel) * D * D + e) =
        make_float4(static_cast<float>(a0 * g), static_cast<float>(a1 * g), static_cast<float>(a2 * g),
                    static_cast<float>(a3 * g));
    if (quarter == 0 && tid < D) {
        double a = 0.0;
        for (int c = c0; c < c1; ++c) a += drel_part[static_cast<size_t>(c) * D + tid];
        d_rel[static_cast<size_t>(rel) * D + tid] = static_cast<float>(a * g);
    }
}

// scatter: one warp per sorted position that starts a run of equal nodes
__global__ void __launch_bounds__(256)
transr_scatter_kernel(const int32_t* __restrict__ ent_node, const int32_t* __restrict__ ent_row, int P,
                      const float* __restrict__ dE_rows, const float* __restrict__ grad,
                      float* __restrict__ d_entity) {
    const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (p >= P) return;
    const int node = ent_node[p];
    if (p > 0 && ent_node[p - 1] == node) return;
    const float g = grad ? *grad : 1.0f;
    float2 acc = make_float2(0.f, 0.f);
    for (int q = p; q < P && ent_node[q] == node; ++q) {
        const float2 v = *reinterpret_cast<const float2*>(dE_rows + static_cast<size_t>(ent_row[q]) * D + lane * 2);
        acc.x += v.x;
        acc.y += v.y;
    }
    *reinterpret_cast<float2*>(d_entity + static_cast<size_t>(node) * D + lane * 2) = make_float2(acc.x * g, acc.y * g);
}

int next_pow2(int64_t n) {
    int p = 1;
    while (p < n) p <<= 1;
    return p;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" size_t kgat_transr_workspace_bytes(int64_t batch, int64_t n_rel) {
    if (batch <= 0 || n_rel <= 0) return 0;
    return carve(nullptr, batch, n_rel).bytes;
}

static int transr_check(int64_t n_nodes, int64_t n_rel, int d, int rd, int64_t batch) {
    KGAT_REQUIRE(n_nodes > 0 && n_rel > 0 && batch > 0);
    if (d != D || rd != D || n_rel > MAX_REL || 3 * batch > MAX_ROWS) return KGAT_ERR_UNSUPPORTED_DIM;
    return KGAT_OK;
}

extern "C" int kgat_transr_fwd(const float* entity, int64_t n_nodes, const float* rel_emb, const float* W_R,
                               int64_t n_rel, int d, int rd, const int64_t* h, const int64_t* r,
                               const int64_t* pos_t, const int64_t* neg_t, int64_t batch, float reg_lambda,
                               float* loss, void* workspace, size_t workspace_bytes, kgat_stream_t stream) {
    KGAT_REQUIRE(entity && rel_emb && W_R && h && r && pos_t && neg_t && loss);
    KGAT_REQUIRE(aligned16(entity) && aligned16(rel_emb) && aligned16(W_R) && aligned16(workspace));
    if (const int rc = transr_check(n_nodes, n_rel, d, rd, batch)) return rc;
    const Ws w = carve(workspace, batch, n_rel);
    if (!workspace || workspace_bytes < w.bytes) return KGAT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int B = static_cast<int>(batch), R = static_cast<int>(n_rel), P = 3 * B;
    const int n2_ent = next_pow2(P) < PLAN_THREADS ? PLAN_THREADS : next_pow2(P);
    int idx_bits = 0;
    while ((1 << idx_bits) < n2_ent) ++idx_bits;
    const bool narrow = n_nodes < (static_cast<int64_t>(1) << (32 - idx_bits));      // (node, row) fits 32 bits
    const size_t key_bytes = static_cast<size_t>(n2_ent) * (narrow ? 4 : 8);
    const size_t cnt_bytes = sizeof(int) * PLAN_WARPS * static_cast<size_t>(R);
    const size_t plan_smem = key_bytes > cnt_bytes ? key_bytes : cnt_bytes;

    static PerDeviceOnce configured;
    if (configured.need()) {
        const int plan_max = MAX_ROWS * (int)sizeof(unsigned long long);
        KGAT_CUDA(cudaFuncSetAttribute(transr_plan_kernel<uint32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       plan_max));
        KGAT_CUDA(cudaFuncSetAttribute(transr_plan_kernel<unsigned long long>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, plan_max));
        KGAT_CUDA(cudaFuncSetAttribute(transr_main_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(MainSmem)));
        configured.set();
    }
    if (narrow)
        transr_plan_kernel<uint32_t><<<1, PLAN_THREADS, plan_smem, st>>>(
            h, r, pos_t, neg_t, B, R, n2_ent, idx_bits, w.order_rel, w.chunk_rel, w.chunk_start, w.chunk_cnt,
            w.rel_chunk0, w.n_chunks, w.done, w.ent_node, w.ent_row);
    else
        transr_plan_kernel<unsigned long long><<<1, PLAN_THREADS, plan_smem, st>>>(
            h, r, pos_t, neg_t, B, R, n2_ent, idx_bits, w.order_rel, w.chunk_rel, w.chunk_start, w.chunk_cnt,
            w.rel_chunk0, w.n_chunks, w.done, w.ent_node, w.ent_row);
    KGAT_LAUNCH_CHECK();
    const int max_chunks = B / T + R;
    transr_main_kernel<<<max_chunks, THREADS, sizeof(MainSmem), st>>>(
        entity, rel_emb, W_R, h, pos_t, neg_t, B, reg_lambda, w.order_rel, w.chunk_rel, w.chunk_start, w.chunk_cnt,
        w.n_chunks, w.done, w.dE_rows, w.dW_part, w.drel_part, w.loss_part, loss);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_transr_bwd(const float* grad_loss, int64_t n_nodes, int64_t n_rel, int d, int rd, int64_t batch,
                               float* d_entity, float* d_rel, float* d_W, const void* workspace,
                               size_t workspace_bytes, kgat_stream_t stream) {
    KGAT_REQUIRE(d_entity && d_rel && d_W);
    KGAT_REQUIRE(aligned16(d_entity) && aligned16(d_rel) && aligned16(d_W) && aligned16(workspace));
    if (const int rc = transr_check(n_nodes, n_rel, d, rd, batch)) return rc;
    const Ws w = carve(const_cast<void*>(workspace), batch, n_rel);
    if (!workspace || workspace_bytes < w.bytes) return KGAT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int R = static_cast<int>(n_rel), P = 3 * static_cast<int>(batch);
    KGAT_CUDA(cudaMemsetAsync(d_entity, 0, sizeof(float) * static_cast<size_t>(n_nodes) * D, st));
    transr_reduce_kernel<<<dim3(R, 4), 256, 0, st>>>(w.rel_chunk0, w.dW_part, w.drel_part, grad_loss, d_W, d_rel);
    KGAT_LAUNCH_CHECK();
    transr_scatter_kernel<<<(P + 7) / 8, 256, 0, st>>>(w.ent_node, w.ent_row, P, w.dE_rows, grad_loss, d_entity);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
