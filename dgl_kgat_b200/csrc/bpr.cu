// BPR loss head of the CF step, fused (reference models.py:170-178):
//   loss = -mean(logsigmoid(<s,p> - <s,n>)) + lambda * (mean|s|^2/2 + mean|p|^2/2 + mean|n|^2/2)
// with s, p, n = embedding[src], embedding[pos], embedding[neg].
//
// Forward: one warp per batch element gathers its three rows with 128-bit loads; a fixed-order
// two-stage fp64 reduction produces the scalar.  Backward: the gradient touches at most 3B of the N
// embedding rows, so instead of a dense (N, dim) tensor it is produced COMPACT: the slots that name the
// same node are chained through an N-sized head array (one integer atomicExch per slot), the smallest
// slot of a chain is the node's leader and its warp sums the chain's contributions in ascending slot
// order (deterministic, no floating-point atomics, no sort) into rows[leader], and row_slot[node] points
// there (-1 elsewhere).  kgat_layer_bwd / kgat_input_bwd read the gradient through row_slot.
// Three graph nodes (memset, link, accumulate) where the first version spent seven (a CUB radix sort of the
// 3B keys among them) - 0.08 ms of a 1.36 ms step.
#include <limits.h>

#include "common.cuh"

namespace kgat {
namespace {

constexpr int BPR_THREADS = 256;
constexpr int BPR_WARPS = BPR_THREADS / 32;

__device__ __forceinline__ float dot4(const float4& a, const float4& b) {
    return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}

__global__ void __launch_bounds__(BPR_THREADS)
bpr_fwd_kernel(const float* __restrict__ emb, int64_t stride, int dim, const int64_t* __restrict__ src,
               const int64_t* __restrict__ pos, const int64_t* __restrict__ neg, int B,
               float* __restrict__ sig, double* __restrict__ partial) {
    __shared__ double s_cf[BPR_WARPS], s_reg[BPR_WARPS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BPR_WARPS + warp;
    double cf = 0.0, reg = 0.0;
    if (b < B) {
        const float* s = emb + src[b] * stride;
        const float* p = emb + pos[b] * stride;
        const float* n = emb + neg[b] * stride;
        float sp = 0.f, sn = 0.f, ss = 0.f, pp = 0.f, nn = 0.f;
        for (int c = lane * 4; c < dim; c += 128) {
            const float4 a = ldg4(s + c), q = ldg4(p + c), r = ldg4(n + c);
            sp += dot4(a, q); sn += dot4(a, r); ss += dot4(a, a); pp += dot4(q, q); nn += dot4(r, r);
        }
        sp = warp_sum(sp); sn = warp_sum(sn); ss = warp_sum(ss); pp = warp_sum(pp); nn = warp_sum(nn);
        const float x = sp - sn;
        // -logsigmoid(x) = -(min(x, 0) - log1p(exp(-|x|)))
        cf = -static_cast<double>(fminf(x, 0.f) - log1pf(expf(-fabsf(x))));
        reg = 0.5 * (static_cast<double>(ss) + static_cast<double>(pp) + static_cast<double>(nn));
        if (lane == 0) sig[b] = 1.0f / (1.0f + expf(x));            // sigmoid(-x)
    }
    if (lane == 0) { s_cf[warp] = cf; s_reg[warp] = reg; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, r = 0.0;
        for (int w = 0; w < BPR_WARPS; ++w) { a += s_cf[w]; r += s_reg[w]; }
        partial[2 * blockIdx.x] = a;
        partial[2 * blockIdx.x + 1] = r;
    }
}

__global__ void __launch_bounds__(256)
bpr_finish_kernel(const double* __restrict__ partial, int n_blocks, int B, float lambda, float* __restrict__ loss) {
    __shared__ double s_a[256], s_r[256];
    double a = 0.0, r = 0.0;
    for (int i = threadIdx.x; i < n_blocks; i += 256) { a += partial[2 * i]; r += partial[2 * i + 1]; }
    s_a[threadIdx.x] = a; s_r[threadIdx.x] = r;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { s_a[threadIdx.x] += s_a[threadIdx.x + o]; s_r[threadIdx.x] += s_r[threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) loss[0] = static_cast<float>(s_a[0] / B + static_cast<double>(lambda) * s_r[0] / B);
}

// Backward, step 1: thread per slot i (slot = role * B + b, role 0/1/2 = src/pos/neg).  The slots that
// name the same node are chained through head[node] (= row_slot, -1 everywhere on entry):
// next[i] = previous head.  The chain ORDER depends on the arrival order of the atomics and is not used
// for anything but membership.  Slots of nodes this rank does not own are marked -2.
__global__ void bpr_link_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ pos,
                                const int64_t* __restrict__ neg, int B, int64_t own_begin, int64_t own_end,
                                int32_t* __restrict__ head, int32_t* __restrict__ next) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 3 * B) return;
    const int role = i / B, b = i - role * B;
    const int64_t id = role == 0 ? src[b] : (role == 1 ? pos[b] : neg[b]);
    next[i] = (id >= own_begin && id < own_end) ? atomicExch(head + id, i) : -2;
}

// Backward, step 2: one warp per slot.  The warp walks the chain of its node; the slot that is the
// SMALLEST member is the node's leader (exactly one per node, the same one in every run) and sums the
// contributions of all members in ascending slot order - a fixed order, so the gradient is bit-identical
// from run to run without a sort and without floating-point atomics - into rows[leader slot];
// row_slot[node] then points there (the chain head it held is no longer needed: only this warp read it).
template <int NG>       // float4 column groups per lane: dim <= 128 * NG
__global__ void __launch_bounds__(BPR_THREADS)
bpr_bwd_kernel(const float* __restrict__ emb, int64_t stride, int dim, const int64_t* __restrict__ src,
               const int64_t* __restrict__ pos, const int64_t* __restrict__ neg,
               const int64_t* __restrict__ node_src, const int64_t* __restrict__ node_pos,
               const int64_t* __restrict__ node_neg, int B, float lambda,
               const float* __restrict__ sig, const float* __restrict__ grad_loss,
               const int32_t* __restrict__ next, float* __restrict__ rows, int64_t rows_stride,
               int32_t* __restrict__ row_slot, int32_t* __restrict__ rows_node) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * BPR_WARPS + warp;
    if (i >= 3 * B) return;
    const int nx = next[i];
    if (nx == -2) {                                  // node owned by another rank
        if (lane == 0) rows_node[i] = -1;
        return;
    }
    const int my_role = i / B, my_b = i - my_role * B;
    const int64_t node = my_role == 0 ? node_src[my_b] : (my_role == 1 ? node_pos[my_b] : node_neg[my_b]);
    // One walk over the chain: member t goes to lane t % 32, register t / 32 (chains of up to 256 slots
    // live in registers; a longer one - a node drawn hundreds of times in one batch - falls back to
    // re-walking the chain for every member, still correct).
    constexpr int PER = 8;
    const int first = *reinterpret_cast<const volatile int32_t*>(row_slot + node);   // chain head, read ONCE
    int m[PER];
#pragma unroll
    for (int r = 0; r < PER; ++r) m[r] = INT_MAX;
    int smallest = INT_MAX, len = 0;
    for (int j = first; j >= 0; j = next[j]) {
        smallest = min(smallest, j);
#pragma unroll
        for (int r = 0; r < PER; ++r)
            if ((len >> 5) == r && (len & 31) == lane) m[r] = j;
        ++len;
    }
    if (smallest != i) {
        if (lane == 0) rows_node[i] = -1;
        return;
    }
    const bool in_regs = len <= 32 * PER;
    const float scale = grad_loss[0] / static_cast<float>(B);
    // each lane owns up to NG float4 column groups: columns lane * 4 + 128 k
    float4 acc[NG];
#pragma unroll
    for (int k = 0; k < NG; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int c0 = lane * 4;
    int done = -1;                                   // largest slot already accumulated
    for (int t = 0; t < len; ++t) {
        // next member in ascending order: the smallest slot above `done`
        int slot = INT_MAX;
        if (in_regs) {
#pragma unroll
            for (int r = 0; r < PER; ++r)
                if (m[r] > done && m[r] < slot) slot = m[r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) slot = min(slot, __shfl_xor_sync(FULL_MASK, slot, o));
        } else {
            for (int j = first; j >= 0; j = next[j])
                if (j > done && j < slot) slot = j;
        }
        done = slot;
        const int role = slot / B, b = slot - role * B;
        const float sg = __ldg(sig + b);
        const float* s = emb + src[b] * stride;
        const float* p = emb + pos[b] * stride;
        const float* n = emb + neg[b] * stride;
        // d loss / d s = scale * (-sg * (p - n) + lambda * s);  d/dp = scale * (-sg * s + lambda * p);
        // d/dn = scale * (sg * s + lambda * n)
        auto contrib = [&](int c, float4& acc) {
            if (c >= dim) return;
            const float4 a = ldg4(s + c);
            float4 g;
            if (role == 0) {
                const float4 q = ldg4(p + c), r = ldg4(n + c);
                g = make_float4(fmaf(-sg, q.x - r.x, lambda * a.x), fmaf(-sg, q.y - r.y, lambda * a.y),
                                fmaf(-sg, q.z - r.z, lambda * a.z), fmaf(-sg, q.w - r.w, lambda * a.w));
            } else if (role == 1) {
                const float4 q = ldg4(p + c);
                g = make_float4(fmaf(-sg, a.x, lambda * q.x), fmaf(-sg, a.y, lambda * q.y),
                                fmaf(-sg, a.z, lambda * q.z), fmaf(-sg, a.w, lambda * q.w));
            } else {
                const float4 r = ldg4(n + c);
                g = make_float4(fmaf(sg, a.x, lambda * r.x), fmaf(sg, a.y, lambda * r.y),
                                fmaf(sg, a.z, lambda * r.z), fmaf(sg, a.w, lambda * r.w));
            }
            acc.x = fmaf(scale, g.x, acc.x); acc.y = fmaf(scale, g.y, acc.y);
            acc.z = fmaf(scale, g.z, acc.z); acc.w = fmaf(scale, g.w, acc.w);
        };
#pragma unroll
        for (int k = 0; k < NG; ++k) contrib(c0 + 128 * k, acc[k]);
    }
    float* out = rows + static_cast<int64_t>(i) * rows_stride;
#pragma unroll
    for (int k = 0; k < NG; ++k)
        if (c0 + 128 * k < dim) *reinterpret_cast<float4*>(out + c0 + 128 * k) = acc[k];
    __syncwarp();                                    // every lane has finished reading the chain head
    if (lane == 0) {
        row_slot[node] = i;
        rows_node[i] = static_cast<int32_t>(node);
    }
}

inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" size_t kgat_bpr_workspace_bytes(int64_t batch) {
    if (batch <= 0) return 256;
    const size_t fwd = align_up(static_cast<size_t>(2) * sizeof(double) * ((batch + BPR_WARPS - 1) / BPR_WARPS));
    const size_t bwd = align_up(static_cast<size_t>(3) * sizeof(int32_t) * batch);       // the slot chains
    return fwd > bwd ? fwd : bwd;
}

extern "C" int kgat_bpr_loss_fwd(const float* emb, int64_t emb_stride, int dim, const int64_t* src,
                                 const int64_t* pos, const int64_t* neg, int64_t batch, float reg_lambda,
                                 float* loss, float* sig, void* workspace, size_t workspace_bytes,
                                 kgat_stream_t stream) {
    KGAT_REQUIRE(emb && src && pos && neg && loss && sig && workspace);
    KGAT_REQUIRE(batch > 0 && batch < (int64_t(1) << 28) && aligned16(emb) && emb_stride % 4 == 0);
    if (dim <= 0 || dim % 4 != 0 || dim > 512) return KGAT_ERR_UNSUPPORTED_DIM;
    if (workspace_bytes < kgat_bpr_workspace_bytes(batch)) return KGAT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int B = static_cast<int>(batch);
    const int blocks = (B + BPR_WARPS - 1) / BPR_WARPS;
    double* partial = static_cast<double*>(workspace);
    bpr_fwd_kernel<<<blocks, BPR_THREADS, 0, st>>>(emb, emb_stride, dim, src, pos, neg, B, sig, partial);
    KGAT_LAUNCH_CHECK();
    bpr_finish_kernel<<<1, 256, 0, st>>>(partial, blocks, B, reg_lambda, loss);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_bpr_loss_bwd(const float* emb, int64_t emb_stride, int dim, const int64_t* src,
                                 const int64_t* pos, const int64_t* neg, int64_t batch, float reg_lambda,
                                 const float* sig, const float* grad_loss,
                                 const int64_t* node_src, const int64_t* node_pos, const int64_t* node_neg,
                                 int64_t own_begin, int64_t own_end, int64_t n_nodes,
                                 float* rows, int32_t* rows_node, int32_t* row_slot,
                                 void* workspace, size_t workspace_bytes, kgat_stream_t stream) {
    KGAT_REQUIRE((node_src == nullptr) == (node_pos == nullptr) && (node_src == nullptr) == (node_neg == nullptr));
    KGAT_REQUIRE(own_begin >= 0 && own_end <= n_nodes && own_begin <= own_end && n_nodes < (int64_t(1) << 31));
    KGAT_REQUIRE(emb && src && pos && neg && sig && grad_loss && rows && rows_node && row_slot && workspace);
    KGAT_REQUIRE(batch > 0 && batch < (int64_t(1) << 28) && n_nodes > 0 && aligned16(emb) && aligned16(rows) &&
                 emb_stride % 4 == 0);
    if (dim <= 0 || dim % 4 != 0 || dim > 512) return KGAT_ERR_UNSUPPORTED_DIM;
    if (workspace_bytes < kgat_bpr_workspace_bytes(batch)) return KGAT_ERR_WORKSPACE;
    cudaStream_t st = as_stream(stream);
    const int B = static_cast<int>(batch);
    int32_t* next = static_cast<int32_t*>(workspace);
    const int64_t* ns = node_src ? node_src : src;
    const int64_t* np_ = node_pos ? node_pos : pos;
    const int64_t* nn = node_neg ? node_neg : neg;
    KGAT_CUDA(cudaMemsetAsync(row_slot, 0xff, sizeof(int32_t) * n_nodes, st));          // -1 everywhere: empty chains
    bpr_link_kernel<<<(3 * B + 255) / 256, 256, 0, st>>>(ns, np_, nn, B, own_begin, own_end, row_slot, next);
    KGAT_LAUNCH_CHECK();
    const unsigned grid = (3 * B + BPR_WARPS - 1) / BPR_WARPS;
    if (dim <= 256)
        bpr_bwd_kernel<2><<<grid, BPR_THREADS, 0, st>>>(emb, emb_stride, dim, src, pos, neg, ns, np_, nn, B, reg_lambda, sig,
                                                      grad_loss, next, rows, dim, row_slot, rows_node);
    else        // 128-wide models: the concat of three layers is 352 floats
        bpr_bwd_kernel<4><<<grid, BPR_THREADS, 0, st>>>(emb, emb_stride, dim, src, pos, neg, ns, np_, nn, B, reg_lambda, sig,
                                                      grad_loss, next, rows, dim, row_slot, rows_node);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
