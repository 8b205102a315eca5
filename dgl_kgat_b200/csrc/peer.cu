// (8) Peer-memory exchange of the destination-partitioned step: arenas, the device-side barrier and the
// small push kernels that have no single-GPU counterpart (see peer.cuh for the scheme; the row pushes of
// the layer / softmax kernels live next to those kernels in propagate.cu and attention.cu).
//
// Barrier: every arena starts with a header of 32-bit words
//   word 0..7   flag[k]  = number of the last barrier rank k has entered (stored by rank k over NVLink)
//   word 32     the owner's barrier count        word 64     the owner's error flag
// Barrier b on rank r: count = b; fence.sys; st.release.sys flag[r] = b in every peer's header; spin with
// ld.acquire.sys on the own header until every flag[k] >= b.  Counts only grow, so a fast rank that is
// already in barrier b+1 cannot be mistaken for a late one and nothing is ever reset.  The release is
// cumulative over everything the stream executed before the barrier kernel (the row pushes of the
// preceding kernels), the acquire orders the kernels that follow behind the peers' pushes.
#include <cstring>

#include "peer.cuh"

namespace kgat {
namespace {

constexpr int EPOCH_WORD = 32;
constexpr int ERROR_WORD = 64;
static_assert((ERROR_WORD + 1) * 4 <= KGAT_PEER_HEADER_BYTES, "header too small");
static_assert(sizeof(cudaIpcMemHandle_t) == KGAT_IPC_HANDLE_BYTES, "cudaIpcMemHandle_t size changed");

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

struct BarrierArgs {
    int world, rank;
    unsigned* hdr[KGAT_MAX_RANKS];
    unsigned long long timeout_ns;
};

__global__ void __launch_bounds__(32) peer_barrier_kernel(BarrierArgs a) {
    unsigned* own = a.hdr[a.rank];
    __shared__ unsigned s_count;
    if (threadIdx.x == 0) {
        const unsigned c = own[EPOCH_WORD] + 1;
        own[EPOCH_WORD] = c;
        s_count = c;
    }
    __syncwarp();
    const unsigned count = s_count;
    const int k = threadIdx.x;
    if (k < a.world && k != a.rank) {
        __threadfence_system();
        st_release_sys(a.hdr[k] + a.rank, count);                 // always signal, even after an error
        if (*reinterpret_cast<volatile unsigned*>(own + ERROR_WORD) == 0) {
            const unsigned long long t0 = global_ns();
            while (static_cast<int>(ld_acquire_sys(own + k) - count) < 0) {
                if (global_ns() - t0 > a.timeout_ns) {
                    *reinterpret_cast<volatile unsigned*>(own + ERROR_WORD) = 1u + static_cast<unsigned>(k);
                    break;
                }
                __nanosleep(32);
            }
        }
    }
    __syncwarp();
}

// one warp per slot; only the owner of the slot's node writes (to every rank's table)
__global__ void __launch_bounds__(256)
gather_rows_push_kernel(const float* __restrict__ emb, int64_t stride, int dim, const int64_t* __restrict__ ids,
                        int n_slots, int64_t own_begin, int64_t own_end, float* __restrict__ table, PeerOut po) {
    const int s = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= n_slots) return;
    const int64_t node = ids[s];
    if (node < own_begin || node >= own_end) return;
    const float* src = emb + node * stride;
    const size_t o = static_cast<size_t>(s) * dim;
    for (int c = lane * 4; c < dim; c += 128) {
        const float4 v = ldg4(src + c);
        *reinterpret_cast<float4*>(table + o + c) = v;
        peer_store(po, o + c, v);
    }
}

__global__ void push_slot_kernel(const float* __restrict__ src, int64_t n, float* __restrict__ own_slot, PeerOut po) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float v = src[i];
        own_slot[i] = v;
        peer_store(po, static_cast<size_t>(i), v);
    }
}

// dst[i] = src[i] in the own arena and in every peer's (n4 float4 elements, grid-stride): the plain all-gather
// push of a contiguous buffer, and the kernel profiles/push_bw.py times to see what one rank can push
__global__ void __launch_bounds__(256)
push_buffer_kernel(const float4* __restrict__ src, int64_t n4, float4* __restrict__ dst, PeerOut po, int local) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float4 v = __ldg(src + i);
        if (local) dst[i] = v;
        peer_store(po, static_cast<size_t>(i) * 4, v);
    }
}

__global__ void sum_slots_kernel(const float* __restrict__ slots, int64_t n, int world, float* __restrict__ out) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float t = slots[i];
    for (int k = 1; k < world; ++k) t += slots[k * n + i];        // rank order: identical bits on every rank
    out[i] = t;
}

}  // namespace
}  // namespace kgat

using namespace kgat;

extern "C" int kgat_peer_arena_create(size_t bytes, void** arena_out_host, void* handle_out_host) {
    KGAT_REQUIRE(arena_out_host && handle_out_host && bytes >= KGAT_PEER_HEADER_BYTES);
    void* p = nullptr;
    KGAT_CUDA(cudaMalloc(&p, bytes));
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return cuda_fail(e);
    }
    memcpy(handle_out_host, &h, sizeof(h));
    *arena_out_host = p;
    return KGAT_OK;
}

extern "C" int kgat_peer_arena_open(const void* handle_host, void** arena_out_host) {
    KGAT_REQUIRE(handle_host && arena_out_host);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_host, sizeof(h));
    void* p = nullptr;
    KGAT_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *arena_out_host = p;
    return KGAT_OK;
}

extern "C" int kgat_peer_arena_close(void* mapped_arena) {
    KGAT_REQUIRE(mapped_arena);
    KGAT_CUDA(cudaIpcCloseMemHandle(mapped_arena));
    return KGAT_OK;
}

extern "C" int kgat_peer_arena_destroy(void* own_arena) {
    KGAT_REQUIRE(own_arena);
    KGAT_CUDA(cudaFree(own_arena));
    return KGAT_OK;
}

extern "C" int kgat_peer_barrier(const kgat_peer_group_t* pg, int timeout_ms, kgat_stream_t stream) {
    KGAT_REQUIRE(peer_group_ok(pg));
    BarrierArgs a;
    a.world = pg->world;
    a.rank = pg->rank;
    for (int k = 0; k < KGAT_MAX_RANKS; ++k) a.hdr[k] = k < pg->world ? static_cast<unsigned*>(pg->arena[k]) : nullptr;
    a.timeout_ns = static_cast<unsigned long long>(timeout_ms > 0 ? timeout_ms : 20000) * 1000000ull;
    peer_barrier_kernel<<<1, 32, 0, as_stream(stream)>>>(a);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_peer_error(const kgat_peer_group_t* pg) {
    KGAT_REQUIRE(peer_group_ok(pg));
    unsigned v = 0;
    KGAT_CUDA(cudaDeviceSynchronize());
    KGAT_CUDA(cudaMemcpy(&v, static_cast<const unsigned*>(pg->arena[pg->rank]) + ERROR_WORD, sizeof(v),
                         cudaMemcpyDeviceToHost));
    return static_cast<int>(v);
}

extern "C" int kgat_peer_gather_rows_push(const float* emb, int64_t emb_stride, int dim, const int64_t* ids,
                                          int64_t n_slots, int64_t own_begin, int64_t own_end, float* table,
                                          const kgat_peer_group_t* pg, kgat_stream_t stream) {
    KGAT_REQUIRE(emb && ids && table && n_slots >= 0 && n_slots < (1ll << 31) && dim > 0 && dim % 4 == 0);
    KGAT_REQUIRE(aligned16(emb) && aligned16(table) && emb_stride % 4 == 0);
    PeerOut po;
    KGAT_REQUIRE(peer_aliases(pg, table, sizeof(float) * static_cast<size_t>(n_slots) * dim, &po));
    if (n_slots == 0) return KGAT_OK;
    const unsigned grid = static_cast<unsigned>((n_slots + 7) / 8);
    gather_rows_push_kernel<<<grid, 256, 0, as_stream(stream)>>>(emb, emb_stride, dim, ids, static_cast<int>(n_slots),
                                                                  own_begin, own_end, table, po);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_peer_push_slot(const float* src, int64_t n, float* slots, const kgat_peer_group_t* pg,
                                   kgat_stream_t stream) {
    KGAT_REQUIRE(src && slots && n >= 0 && peer_group_ok(pg));
    float* own_slot = slots + static_cast<size_t>(pg->rank) * n;
    PeerOut po;
    KGAT_REQUIRE(peer_aliases(pg, own_slot, sizeof(float) * static_cast<size_t>(n), &po));
    // the whole (world, n) buffer must lie in the arena, not just the own slot
    PeerOut whole;
    KGAT_REQUIRE(peer_aliases(pg, slots, sizeof(float) * static_cast<size_t>(n) * pg->world, &whole));
    if (n == 0) return KGAT_OK;
    const unsigned grid = static_cast<unsigned>(n + 255) / 256 < 64u ? static_cast<unsigned>(n + 255) / 256 : 64u;
    push_slot_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, n, own_slot, po);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_peer_push_buffer(const float* src, int64_t n, float* dst, const kgat_peer_group_t* pg,
                                     kgat_stream_t stream) {
    KGAT_REQUIRE(src && dst && n >= 0 && n % 4 == 0 && aligned16(src) && aligned16(dst) && peer_group_ok(pg));
    PeerOut po;
    KGAT_REQUIRE(peer_aliases(pg, dst, sizeof(float) * static_cast<size_t>(n), &po));
    if (n == 0) return KGAT_OK;
    const int64_t n4 = n / 4, want = (n4 + 255) / 256, cap = 8ll * num_sms();
    // with the multicast mapping the switch writes the own copy too: no separate local store
    push_buffer_kernel<<<static_cast<unsigned>(want < cap ? want : cap), 256, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(src), n4, reinterpret_cast<float4*>(dst), po, po.mc ? 0 : 1);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}

extern "C" int kgat_peer_sum_slots(const float* slots, int64_t n, int world, float* out, kgat_stream_t stream) {
    KGAT_REQUIRE(slots && out && n >= 0 && world >= 1 && world <= KGAT_MAX_RANKS);
    if (n == 0) return KGAT_OK;
    sum_slots_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, as_stream(stream)>>>(slots, n, world, out);
    KGAT_LAUNCH_CHECK();
    return KGAT_OK;
}
