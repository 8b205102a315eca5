// Peer-memory exchange for the destination-partitioned (multi-GPU) step.
//
// Every rank owns one ARENA (a cudaMalloc'ed block exported with cudaIpc) and maps the arenas of the
// other ranks of the box into its own address space (NVLink / NVSwitch peer access).  A buffer that has to
// reach every rank (a layer output, dh_nb, the entity gradient, the attention weights) lives at the SAME
// offset in every arena, so the kernel that produces a row stores it into its own arena and, in the same
// instruction stream, into the arena of every peer: the all-gather is fused into the producing kernel and
// overlaps its math row by row.  What is left of the collective is a barrier (peer.cu).
#pragma once

#include "common.cuh"

namespace kgat {

// Remote aliases of ONE output buffer: base[k] is the address of that buffer inside the k-th peer's arena
// (peers only, the own copy is written through the kernel's normal output pointer).  Passed by value.
struct PeerOut {
    int n;
    float* base[KGAT_MAX_RANKS - 1];
};

#ifdef __CUDACC__
// store v at float offset `off` of the buffer in every peer's arena (constant indices: the table stays
// in the kernel's parameter bank)
template <typename T>
__device__ __forceinline__ void peer_store(const PeerOut& po, size_t off, const T& v) {
#pragma unroll
    for (int k = 0; k < KGAT_MAX_RANKS - 1; ++k)
        if (k < po.n) *reinterpret_cast<T*>(po.base[k] + off) = v;
}
#endif

inline PeerOut no_peers() {
    PeerOut p;
    p.n = 0;
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
    return true;
}

}  // namespace kgat
