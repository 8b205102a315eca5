// Hardware self-test of the tcgen05 building blocks in umma.cuh: C[128 x N] = A[128 x K] * B[N x K]^T
// on ONE CTA, operands staged in the K-major no-swizzle core-matrix layout, accumulator in TMEM.
//   mode 0: one kind::tf32 pass (expected error ~1e-3 relative)
//   mode 1: 3xTF32 (hi*hi + hi*lo + lo*hi), expected ~fp32 accuracy
//   variant 0: no-swizzle layout, descriptor LBO = K-direction core stride, SBO = 8-row group stride (the
//              documented meaning); variant 1: the two swapped (a negative control: must NOT reproduce
//              the GEMM); variant 2: SWIZZLE_128B layout, K = 32 or 64 held as K-halves of 128-byte rows
// A wrong descriptor cannot hang the GPU: the completion wait is bounded and reports through *err.
#include "common.cuh"
#include "umma.cuh"

namespace kgat {
namespace {

__global__ void __launch_bounds__(128)
umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int N, int K,
                     int variant, int mode, int* __restrict__ err) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    unsigned char* sAhi = smem;
    unsigned char* sAlo = sAhi + 128 * K * 4;
    unsigned char* sBhi = sAlo + 128 * K * 4;
    unsigned char* sBlo = sBhi + N * K * 4;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // variant 2: each operand is a sequence of K/32 SWIZZLE_128B half-tiles (rows x 128 bytes)
    auto offs = [&](int r, int k, int rows) -> uint32_t {
        return variant == 2 ? static_cast<uint32_t>((k >> 5) * rows * 128) + umma::sw128_offset(r, k & 31)
                            : umma::tile_offset(r, k, K);
    };
    for (int idx = tid; idx < 128 * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        float hi, lo;
        umma::split_tf32(A[idx], hi, lo);
        *reinterpret_cast<float*>(sAhi + offs(r, k, 128)) = hi;
        *reinterpret_cast<float*>(sAlo + offs(r, k, 128)) = lo;
    }
    for (int idx = tid; idx < N * K; idx += 128) {
        const int r = idx / K, k = idx % K;
        float hi, lo;
        umma::split_tf32(B[idx], hi, lo);
        *reinterpret_cast<float*>(sBhi + offs(r, k, N)) = hi;
        *reinterpret_cast<float*>(sBlo + offs(r, k, N)) = lo;
    }
    if (warp == 0) umma::tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        umma::mbar_init(&bar, 1);
        umma::fence_barrier_init();
    }
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = umma::make_idesc_tf32(128, N);
        const uint32_t k_stride = umma::LBO_BYTES, mn_stride = (K / 4) * 128;
        const uint32_t lbo = variant == 0 ? k_stride : mn_stride;
        const uint32_t sbo = variant == 0 ? mn_stride : k_stride;
        for (int ks = 0; ks < K / 8; ++ks) {
            uint64_t ahi, alo, bhi, blo;
            if (variant == 2) {
                const uint32_t offA = (ks >> 2) * 128 * 128 + (ks & 3) * 32, offB = (ks >> 2) * N * 128 + (ks & 3) * 32;
                ahi = umma::make_smem_desc_sw128(umma::smem_u32(sAhi) + offA);
                alo = umma::make_smem_desc_sw128(umma::smem_u32(sAlo) + offA);
                bhi = umma::make_smem_desc_sw128(umma::smem_u32(sBhi) + offB);
                blo = umma::make_smem_desc_sw128(umma::smem_u32(sBlo) + offB);
            } else {
                const uint32_t off = ks * 2 * umma::CORE_BYTES;
                ahi = umma::make_smem_desc(umma::smem_u32(sAhi) + off, lbo, sbo);
                alo = umma::make_smem_desc(umma::smem_u32(sAlo) + off, lbo, sbo);
                bhi = umma::make_smem_desc(umma::smem_u32(sBhi) + off, lbo, sbo);
                blo = umma::make_smem_desc(umma::smem_u32(sBlo) + off, lbo, sbo);
            }
            umma::mma_tf32(tmem, ahi, bhi, idesc, ks > 0);
            if (mode == 1) {
                umma::mma_tf32(tmem, ahi, blo, idesc, 1);
                umma::mma_tf32(tmem, alo, bhi, idesc, 1);
            }
        }
        umma::mma_commit(&bar);
    }
    const bool ok = umma::mbar_wait(&bar, 0);
    umma::fence_after_sync();
    if (!ok) {
        if (tid == 0) *err = 1;
    } else {
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v[32];
            umma::tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) C[static_cast<size_t>(row) * N + c0 + i] = v[i];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" int kgat_selftest_umma(const float* A, const float* B, float* C, int N, int K, int variant, int mode,
                                  int* err_flag, kgat_stream_t stream) {
    KGAT_REQUIRE(A && B && C && err_flag);
    if (N < 32 || N > 256 || N % 32 != 0 || K < 8 || K > 64 || K % 8 != 0) return KGAT_ERR_UNSUPPORTED_DIM;
    if (variant == 2 && K % 32 != 0) return KGAT_ERR_UNSUPPORTED_DIM;
    const size_t smem = static_cast<size_t>(2) * (128 + N) * K * 4;
    KGAT_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(A, B, C, N, K, variant, mode, err_flag);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
