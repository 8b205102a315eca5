// Peer-memory exchange for the destination-partitioned (multi-GPU) step.
//
// Every rank owns one ARENA (a cudaMalloc'ed block exported with cudaIpc) and maps the arenas of the
// other ranks of the box into its own address space (NVLink / NVSwitch peer access).  A buffer that has to
// reach every rank (a layer output, dh_nb, the entity gradient, the attention weights) lives at the SAME
// offset in every arena, so the kernel that produces a row stores it into its own arena and, in the same
// instruction stream, into the arena of every peer: the all-gather is fused into the producing kernel and
// overlaps its math row by row.  What is left of the collective is a barrier (peer.cu).
#pragma once

#include <cstdlib>

#include "common.cuh"

namespace kgat {

// Remote aliases of ONE output buffer: base[k] is the address of that buffer inside the k-th peer's arena
// (peers only, the own copy is written through the kernel's normal output pointer).  Passed by value.
struct PeerOut {
    int n;
    float* base[KGAT_MAX_RANKS - 1];
    float* mc;          // the same buffer inside the NVSwitch multicast mapping of all arenas, or nullptr
};

#ifdef __CUDACC__
// store v at float offset `off` of the buffer in every peer's arena (constant indices: the table stays
// in the kernel's parameter bank)
// One store that the switch replicates into the arena of every rank (the own one included): the row leaves
// this GPU ONCE instead of world - 1 times.  Measured before: a rank's NVLink egress was world - 1 copies of every
// exchanged row (~1.3 KB per row and peer per step, ~430 GB/s effective) and did not shrink with the rank count.
__device__ __forceinline__ void multimem_store(float* p, const float& v) {
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void multimem_store(float* p, const float2& v) {
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void multimem_store(float* p, const float4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

template <typename T>
__device__ __forceinline__ void peer_store(const PeerOut& po, size_t off, const T& v) {
    if (po.mc) {
        multimem_store(po.mc + off, v);
        return;
    }
#pragma unroll
    for (int k = 0; k < KGAT_MAX_RANKS - 1; ++k)
        if (k < po.n) *reinterpret_cast<T*>(po.base[k] + off) = v;
}
#endif

inline PeerOut no_peers() {
    PeerOut p;
    p.n = 0;
    p.mc = nullptr;
    for (int k = 0; k < KGAT_MAX_RANKS - 1; ++k) p.base[k] = nullptr;
    return p;
}

inline bool peer_group_ok(const kgat_peer_group_t* pg) {
    if (!pg || pg->world < 1 || pg->world > KGAT_MAX_RANKS || pg->rank < 0 || pg->rank >= pg->world) return false;
    if (pg->arena_bytes < KGAT_PEER_HEADER_BYTES) return false;
    for (int k = 0; k < pg->world; ++k)
        if (!pg->arena[k]) return false;
    return true;
}

// Translate a local output buffer of `bytes` bytes into its aliases in the peers' arenas.
// Returns false when the buffer does not lie inside the own arena's payload area.
inline bool peer_aliases(const kgat_peer_group_t* pg, const void* local, size_t bytes, PeerOut* out) {
    if (!peer_group_ok(pg) || !local) return false;
    const char* own = static_cast<const char*>(pg->arena[pg->rank]);
    const char* p = static_cast<const char*>(local);
    if (p < own + KGAT_PEER_HEADER_BYTES || p + bytes > own + pg->arena_bytes) return false;
    const size_t off = static_cast<size_t>(p - own);
    *out = no_peers();
    for (int k = 0; k < pg->world; ++k) {
        if (k == pg->rank) continue;
        out->base[out->n++] = reinterpret_cast<float*>(static_cast<char*>(pg->arena[k]) + off);
    }
    if (pg->multicast && pg->world > 1) out->mc = reinterpret_cast<float*>(static_cast<char*>(pg->multicast) + off);
    // MEASUREMENT ONLY (results are wrong): KGAT_DEBUG_NO_REMOTE=1 drops every remote store, so the *_push kernels
    // can be timed without their NVLink traffic (profiles/README.md, what the pushes cost)
    static const bool no_remote = [] { const char* e = std::getenv("KGAT_DEBUG_NO_REMOTE"); return e && e[0] == '1'; }();
    if (no_remote) { out->n = 0; out->mc = nullptr; }
    return true;
}

}  // namespace kgat
