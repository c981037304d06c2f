// Row tiling shared by the per-sample kernels (K1, K2, K3).
//
// Data layout: every image/latent tensor is contiguous [B, D]. A "unit" is one global access of
// W elements of one stream: W = 16/sizeof(T) (one 128-bit access) on the vector path, W = 1 on
// the scalar path taken when D is not a multiple of the vector width or a pointer is not 16-byte
// aligned. A "tile" is `iters * VPT * kThreads` consecutive units of ONE row, processed by one
// CTA; rows longer than a tile are split into `nch` tiles whose partial sums are combined in
// fixed order by the last CTA to finish the row (common.cuh: last_cta_ticket). The host picks
// `iters` so that the grid has a few tiles per resident CTA when the batch is small (so 148 SMs
// stay busy at B = 4) and long tiles when it is large (fewer barriers per byte).
#pragma once

#include "common.cuh"

namespace siss {

constexpr int kMaxRowChunks = 128;  // max tiles per row (workspace: B * 128 * 4 floats)
constexpr int kRowPartialStride = 4;  // floats per (row, chunk) partial slot

struct RowTiling {
    long long B;
    long long D;
    long long units_per_row;  // ceil(D / W); for the vector path D % W == 0
    int iters;                // inner iterations per tile
    int nch;                  // tiles per row
    long long tiles;          // B * nch
    int grid;                 // CTAs to launch (persistent, grid-stride over tiles)
};

int cached_sm_count();

// ctas_per_sm: residency the kernel was compiled for (launch_bounds min blocks).
inline RowTiling make_row_tiling(long long B, long long D, int W, int VPT, int ctas_per_sm) {
    RowTiling rt;
    rt.B = B; rt.D = D;
    rt.units_per_row = (D + W - 1) / W;
    const long long step = (long long)kThreads * VPT;        // units per inner iteration
    const long long max_iters = (rt.units_per_row + step - 1) / step;
    const long long slots = (long long)cached_sm_count() * ctas_per_sm;
    // Largest power-of-two iters (<= 8) that still leaves >= 4 tiles per resident CTA slot.
    long long iters = 8;
    while (iters > 1) {
        long long it = iters < max_iters ? iters : max_iters;
        long long nch = (rt.units_per_row + step * it - 1) / (step * it);
        if (B * nch >= 4 * slots) break;
        iters >>= 1;
    }
    // ...but never more than kMaxRowChunks tiles per row (bounds the partial-sum workspace).
    const long long min_iters = (rt.units_per_row + step * kMaxRowChunks - 1) / (step * kMaxRowChunks);
    if (iters < min_iters) iters = min_iters;
    if (iters > max_iters) iters = max_iters;
    if (iters < 1) iters = 1;
    rt.iters = (int)iters;
    rt.nch = (int)((rt.units_per_row + step * iters - 1) / (step * iters));
    rt.tiles = B * rt.nch;
    rt.grid = (int)(rt.tiles < slots ? rt.tiles : slots);
    if (rt.grid < 1) rt.grid = 1;
    return rt;
}

template <typename T, int W>
__device__ __forceinline__ void load_unit(const T* p, float (&f)[W]) {
    if constexpr (W == 1) {
        f[0] = VecTraits<T>::load1(p);
    } else {
        static_assert(W == VecTraits<T>::N, "vector width");
        VecTraits<T>::unpack(ldg_stream(p), f);
    }
}

template <typename T, int W>
__device__ __forceinline__ void store_unit(T* p, const float (&f)[W]) {
    if constexpr (W == 1) {
        VecTraits<T>::store1(p, f[0]);
    } else {
        stg_stream(p, VecTraits<T>::pack(f));
    }
}

// Raw 128-bit fetch used to batch all loads of an iteration before the first use (MLP).
template <typename T, int W> struct RawUnit { uint4 v; };
template <typename T> struct RawUnit<T, 1> { float v; };

template <typename T, int W>
__device__ __forceinline__ void fetch_raw(const T* p, RawUnit<T, W>& r) {
    if constexpr (W == 1) r.v = VecTraits<T>::load1(p);
    else r.v = ldg_stream(p);
}

template <typename T, int W>
__device__ __forceinline__ void decode_raw(const RawUnit<T, W>& r, float (&f)[W]) {
    if constexpr (W == 1) f[0] = r.v;
    else VecTraits<T>::unpack(r.v, f);
}

// torch-style index: negative wraps once, then clamp for memory safety (eager would raise).
__device__ __forceinline__ int wrap_timestep(long long t, int T) {
    if (t < 0) t += T;
    if (t < 0) t = 0;
    if (t >= T) t = T - 1;
    return (int)t;
}

// Workspace for cross-CTA row reductions: [B] ticket counters (zero between launches; the
// kernels restore that) followed by [B][kMaxRowChunks][kRowPartialStride] fp32 partial slots.
struct RowWorkspace {
    unsigned int* counters;
    float* partials;
};

inline long long row_ws_counter_bytes(long long B) {
    return ((B * (long long)sizeof(unsigned int) + 255) / 256) * 256;
}

inline RowWorkspace carve_row_workspace(void* ws, long long B) {
    RowWorkspace r;
    r.counters = reinterpret_cast<unsigned int*>(ws);
    r.partials = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + row_ws_counter_bytes(B));
    return r;
}

}  // namespace siss
