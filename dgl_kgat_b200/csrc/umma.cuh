// tcgen05 / TMEM / mbarrier building blocks (inline PTX, sm_100a) used by the tensor-core kernels.
//
// Two K-major operand layouts are supported (both verified on hardware by kgat_selftest_umma):
//  * SWIZZLE_128B (used by the kernels): a tile row is 32 fp32 = 128 bytes; rows are 128 B apart,
//    8-row groups 1024 B apart (SBO), and inside a row the eight 16-byte chunks are XOR-permuted by
//    the row index:  byte(r, k) = (r / 8) * 1024 + (r % 8) * 128 + (((k / 4) ^ (r % 8)) * 16) + (k % 4) * 4.
//    Eight consecutive lanes can read one row's 128 contiguous bytes from global memory AND write them
//    as one conflict-free 128-byte shared-memory wavefront.  Tile bases must be 1024-byte aligned.
//    One kind::tf32 MMA consumes K = 8 = 32 bytes: the k-th MMA of a tile starts 32 * k bytes in.
//  * no swizzle ("interleave"): 8-row x 16-byte core matrices,
//    byte(r, k) = (r / 8) * SBO + (k / 4) * 128 + (r % 8) * 16 + (k % 4) * 4, LBO = 128, SBO = (KT / 4) * 128;
//    the k-th MMA starts 256 * k bytes in.
// The accumulator of an M = 128 MMA sits in TMEM with row i in lane i and column j in column j.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace kgat {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: gives up after ~2^31 clocks (about a second) instead of hanging the GPU when a
// descriptor is wrong; the caller reports the failure through an error word.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return true;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > (1ll << 31)) return false;
    }
    return true;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------- TMEM
// executed by ONE full warp; writes the TMEM base address to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive fp32 columns of this thread's TMEM lane (no wait)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- descriptors
constexpr uint32_t CORE_BYTES = 128;          // one 8 x 16B core matrix
constexpr uint32_t LBO_BYTES = CORE_BYTES;    // next core matrix along K

// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (CUTLASS cute::UMMA::SmemDescriptor layout:
// start_address [0,14), leading_byte_offset [16,30), stride_byte_offset [32,46), version [46,48) = 1,
// layout_type [61,64) = 0; all offsets in 16-byte units)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    return d;
}

// K-major SWIZZLE_128B descriptor (layout_type [61,64) = 2): 8-row groups 1024 B apart; the leading
// offset is not used when the MMA's K extent lies inside one 128-byte row (CUTLASS encodes 1)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(1024 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(2) << 61;
    return d;
}

// byte offset of element (r, k), k in [0, 32), inside a SWIZZLE_128B tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int k) {
    return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 2) ^ (r & 7)) & 7) << 4) + (k & 3) * 4);
}

// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major
// (cute::UMMA::InstrDescriptor: c_format [4,6) = 1 (F32), a_format [7,10) = b_format [10,13) = 2 (TF32),
//  a_major bit 15 = b_major bit 16 = 0, n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// arrive on an mbarrier once every MMA issued so far by this thread has completed
// (implies tcgen05.fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- 3xTF32 split
// x = hi + lo with hi the round-to-nearest (ties away) TF32 of x; lo is exact in fp32 and is itself
// truncated to TF32 by the tensor core.  hi*hi + hi*lo + lo*hi (+ lo*lo) recovers ~fp32 accuracy.
// The rounding is done on the bit pattern (add half an ulp of the 10-bit mantissa, clear the low 13
// bits): identical to cvt.rna.tf32.f32 for finite values below the overflow threshold, and 2 integer
// instructions instead of the 5 (add, |x| < inf test, select, mask) ptxas emits for the cvt.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = x - hi;
}

// byte offset of element (r, k) inside a K-major no-swizzle tile whose rows span kt floats
__device__ __forceinline__ uint32_t tile_offset(int r, int k, int kt) {
    return static_cast<uint32_t>((r >> 3) * (kt >> 2) * 128 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

}  // namespace umma
}  // namespace kgat
