// Shared helpers for the sm_100a kernels behind include/kgat_b200.h.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/kgat_b200.h"

namespace kgat {

constexpr unsigned FULL_MASK = 0xffffffffu;
constexpr float NORM_EPS = 1e-12f;   // F.normalize default eps, reference models.py:165

extern thread_local int g_last_cuda_error;

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = static_cast<int>(e);
    return KGAT_ERR_CUDA;
}

#define KGAT_CUDA(expr)                                   \
    do {                                                  \
        cudaError_t _e = (expr);                          \
        if (_e != cudaSuccess) return ::kgat::cuda_fail(_e); \
    } while (0)

#define KGAT_LAUNCH_CHECK() KGAT_CUDA(cudaGetLastError())

#define KGAT_REQUIRE(cond)                     \
    do {                                       \
        if (!(cond)) return KGAT_ERR_INVALID_ARG; \
    } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

inline cudaStream_t as_stream(kgat_stream_t s) { return static_cast<cudaStream_t>(s); }

constexpr int MAX_DEVICES = 64;

inline int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev % MAX_DEVICES;
}

inline int num_sms() {
    static int n[MAX_DEVICES] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        if (cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n[dev] <= 0)
            n[dev] = 148;
    }
    return n[dev];
}

// Function attributes (the opt-in to more than 48 KB of dynamic shared memory) and device-symbol
// addresses are PER DEVICE: a launcher remembers which devices of the process it has prepared.
//     static PerDeviceOnce once;  if (once.need()) { cudaFuncSetAttribute(...); once.set(); }
struct PerDeviceOnce {
    bool done[MAX_DEVICES] = {};
    int cur = 0;
    bool need() { cur = current_device(); return !done[cur]; }
    void set() { done[cur] = true; }
};
template <typename T>
struct PerDevice {
    T v[MAX_DEVICES] = {};
    T& get() { return v[current_device()]; }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

// sum over the lanes of an aligned group of GROUP lanes (GROUP power of two <= 32)
template <int GROUP>
__device__ __forceinline__ float group_sum(float v, unsigned mask = FULL_MASK) {
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
}

// combine the partial float4 accumulators held by the 32/GROUP lane groups of a warp;
// afterwards every group holds the total
template <int GROUP>
__device__ __forceinline__ float4 across_groups_sum(float4 v) {
#pragma unroll
    for (int o = GROUP; o < 32; o <<= 1) {
        v.x += __shfl_xor_sync(FULL_MASK, v.x, o);
        v.y += __shfl_xor_sync(FULL_MASK, v.y, o);
        v.z += __shfl_xor_sync(FULL_MASK, v.z, o);
        v.w += __shfl_xor_sync(FULL_MASK, v.w, o);
    }
    return v;
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}

}  // namespace kgat
