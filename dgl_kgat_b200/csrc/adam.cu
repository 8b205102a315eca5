// (7) Adam step over a list of fp32 tensors in ONE launch.
//
// Replaces torch.optim.Adam.step as the reference uses it (kgat.py:101: Adam(model.parameters(), lr),
// betas (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad), same arithmetic:
//   m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;
//   p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// HBM-bound: 16 bytes read + 12 written per element, 128-bit accesses, every tensor of the list handled
// by the same grid (block -> tensor by a prefix table passed by value).
#include "common.cuh"

#include <math.h>

namespace kgat {
namespace {

constexpr int ADAM_MAX_TENSORS = 16;
constexpr int ADAM_THREADS = 256;
constexpr int ADAM_VEC_PER_BLOCK = ADAM_THREADS * 4;       // float4 per block

struct AdamTable {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long n[ADAM_MAX_TENSORS];
    int block0[ADAM_MAX_TENSORS + 1];
    int count;
};

__global__ void __launch_bounds__(ADAM_THREADS)
adam_kernel(const AdamTable t, float b1, float b2, float c1, float c2, float step_size, float inv_sqrt_bc2, float eps,
            float decay) {
    int k = 0;
    while (k + 1 < t.count && static_cast<int>(blockIdx.x) >= t.block0[k + 1]) ++k;
    const long long n = t.n[k], nv = n >> 2;
    float4* __restrict__ p4 = reinterpret_cast<float4*>(t.p[k]);
    const float4* __restrict__ g4 = reinterpret_cast<const float4*>(t.g[k]);
    float4* __restrict__ m4 = reinterpret_cast<float4*>(t.m[k]);
    float4* __restrict__ v4 = reinterpret_cast<float4*>(t.v[k]);
    const long long base = static_cast<long long>(blockIdx.x - t.block0[k]) * ADAM_VEC_PER_BLOCK;
    auto upd = [&](float& p, float g, float& m, float& v) {
        if (decay != 0.0f) g = fmaf(decay, p, g);
        m = fmaf(b1, m, c1 * g);                       // torch: m.lerp_(g, 1 - b1)
        v = fmaf(b2, v, c2 * g * g);
        const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;
        p -= step_size * (m / denom);
    };
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const long long i = base + u * ADAM_THREADS + threadIdx.x;
        if (i < nv) {
            float4 p = p4[i], m = m4[i], v = v4[i];
            const float4 g = g4[i];
            upd(p.x, g.x, m.x, v.x); upd(p.y, g.y, m.y, v.y); upd(p.z, g.z, m.z, v.z); upd(p.w, g.w, m.w, v.w);
            p4[i] = p; m4[i] = m; v4[i] = v;
        }
    }
    // tail (n not a multiple of 4): last block of the tensor
    if (static_cast<int>(blockIdx.x) + 1 == t.block0[k + 1] && threadIdx.x < (n & 3)) {
        const long long i = (nv << 2) + threadIdx.x;
        upd(t.p[k][i], t.g[k][i], t.m[k][i], t.v[k][i]);
    }
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" int kgat_adam_step(int n_tensors, float* const* params, const float* const* grads, float* const* exp_avg,
                              float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2,
                              double eps, double weight_decay, int64_t step, kgat_stream_t stream) {
    KGAT_REQUIRE(n_tensors >= 0 && params && grads && exp_avg && exp_avg_sq && numel && step >= 1);
    if (n_tensors > ADAM_MAX_TENSORS) return KGAT_ERR_UNSUPPORTED_DIM;
    if (n_tensors == 0) return KGAT_OK;
    AdamTable t;
    t.count = n_tensors;
    int blocks = 0;
    for (int k = 0; k < n_tensors; ++k) {
        KGAT_REQUIRE(params[k] && grads[k] && exp_avg[k] && exp_avg_sq[k] && numel[k] >= 0);
        KGAT_REQUIRE(aligned16(params[k]) && aligned16(grads[k]) && aligned16(exp_avg[k]) && aligned16(exp_avg_sq[k]));
        t.p[k] = params[k]; t.g[k] = grads[k]; t.m[k] = exp_avg[k]; t.v[k] = exp_avg_sq[k]; t.n[k] = numel[k];
        t.block0[k] = blocks;
        const long long nv = (numel[k] + 3) >> 2;
        blocks += static_cast<int>((nv + ADAM_VEC_PER_BLOCK - 1) / ADAM_VEC_PER_BLOCK);
    }
    t.block0[n_tensors] = blocks;
    if (blocks == 0) return KGAT_OK;
    // every scalar is derived in double and rounded once, the way torch derives them from Python floats
    const double bc1 = 1.0 - pow(beta1, static_cast<double>(step));
    const double bc2 = 1.0 - pow(beta2, static_cast<double>(step));
    adam_kernel<<<blocks, ADAM_THREADS, 0, as_stream(stream)>>>(
        t, static_cast<float>(beta1), static_cast<float>(beta2), static_cast<float>(1.0 - beta1),
        static_cast<float>(1.0 - beta2), static_cast<float>(lr / bc1), static_cast<float>(1.0 / sqrt(bc2)),
        static_cast<float>(eps), static_cast<float>(weight_decay));
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
