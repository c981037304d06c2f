// Work decomposition shared by the per-sample kernels (K1, K2, K3).
//
// Data layout: every image/latent tensor is contiguous [B, D]. A "unit" is one global access of
// W elements of one stream: W = 16/sizeof(T) (one 128-bit access) on the vector path, W = 1 on
// the scalar path taken when D is not a multiple of the vector width or a pointer is not 16-byte
// aligned. The flat unit space [0, U), U = B * ceil(D/W), is cut into `grid` CONTIGUOUS spans,
// one per CTA (persistent grid = SMs x resident CTAs, or fewer when there is little work). A CTA
// walks its span row segment by row segment: per-row scalars (t, gamma, sigma, mask, weights) are
// fetched once per segment, per-row sums are accumulated in registers over the whole segment and
// reduced ONCE per segment (warp shuffles -> smem). Long contiguous streams per CTA keep HBM pages
// open; a CTA touches at most span/row_len + 2 rows.
//
// Rows that are split over several CTAs combine their partial sums through a small workspace:
//   a span has at most two row segments shared with other spans (its first and its last row), so it owns
//   two partial slots: slot(span c, segment) = 2c if the row began before the span, 2c + 1 otherwise
//   contributors of row r = owner(first unit of r) .. owner(last unit of r)   (consecutive spans)
// The last contributor to arrive (ticket counter per row) sums the slots in fixed order with one
// warp (lane-strided, then butterfly), so results are bitwise reproducible for a given shape+GPU.
// The layout does not depend on the batch size of the call (see RowWorkspace below).
#pragma once

#include "common.cuh"

namespace siss {

constexpr int kMaxGrid = 148 * 8;         // upper bound on any row-kernel grid
constexpr int kMaxSpans = kMaxGrid * 4;   // upper bound on the number of spans (slots workspace)
constexpr int kRowPartialStride = 4;      // floats per partial slot (one 128-bit access)

struct RowSched {
    long long B;
    long long D;
    long long upr;   // units per row = ceil(D / W)
    long long U;     // total units = B * upr
    int grid;        // CTAs to launch
    int release_all; // TMA ring: 1 = every consumer thread arrives on the stage's empty barrier itself,
                     // 0 = one elected lane per warp arrives after __syncwarp (see bulkpipe.cuh)
    int nspans;      // contiguous spans the unit space is cut into. Currently == grid (span id == blockIdx.x).
                     // A dynamic span queue (several spans per CTA, claimed with atomicAdd) was built and
                     // measured in round 1: at the celeb shape a CTA has only ~14 stages of work, so the
                     // per-span claim + row-setup cost exceeded what balancing recovered (K1oK2 32.6-47 us
                     // vs 31.6 us static); kept out, see DESIGN.md.
};

constexpr int kMaxDevices = 64;           // per-device launch state tables (one process may drive several GPUs)
int current_device_slot();
int cached_sm_count();

// tuning knobs for the dynamically scheduled kernels (read once; see tools/ for the sweep that set the defaults)
int env_int(const char* name, int dflt);

inline RowSched make_row_sched(long long B, long long D, int W, int ctas_per_sm, int spans_per_cta = 1,
                               int min_span_stages = 8) {
    RowSched s;
    s.B = B; s.D = D;
    s.upr = (D + W - 1) / W;
    s.U = B * s.upr;
    long long slots = (long long)cached_sm_count() * ctas_per_sm;
    if (slots > kMaxGrid) slots = kMaxGrid;
    // at least one unit per thread per CTA; small problems use fewer, fully busy CTAs
    long long want = (s.U + kThreads - 1) / kThreads;
    s.grid = (int)(want < slots ? want : slots);
    if (s.grid < 1) s.grid = 1;
    // dynamic scheduling: up to spans_per_cta spans per CTA, but never spans shorter than 8 stages
    long long ns = s.grid;
    if (spans_per_cta > 1) {
        const long long by_size = s.U / ((long long)min_span_stages * kThreads);
        ns = (long long)s.grid * spans_per_cta;
        if (ns > by_size) ns = by_size;
        if (ns < s.grid) ns = s.grid;
        if (ns > kMaxSpans) ns = kMaxSpans;
    }
    s.nspans = (int)ns;
    s.release_all = 0;
    return s;
}

// Oversubscribed launch of the register-staged (LDG) kernels: `k` spans per resident CTA slot, span id == blockIdx.x,
// so the hardware CTA scheduler balances SM-to-SM speed differences (experiment knob SISS_LDG_OVERSUB; spans never
// shorter than `vpt` units per thread).
inline void oversubscribe(RowSched& s, int k, int vpt) {
    if (k <= 1) return;
    long long g = (long long)s.grid * k;
    const long long by_size = s.U / ((long long)kThreads * vpt);
    if (g > by_size) g = by_size;
    if (g > kMaxSpans) g = kMaxSpans;
    if (g > s.grid) { s.grid = (int)g; s.nspans = (int)g; }
}

// span c: [floor(c U / S), floor((c+1) U / S)),  S = nspans
__device__ __forceinline__ void span_range(const RowSched& s, long long c, long long& u0, long long& u1) {
    const long long g = s.nspans;
    u0 = (c * s.U) / g;
    u1 = ((c + 1) * s.U) / g;
}

// statically scheduled kernels: span id == blockIdx.x (nspans == gridDim.x)
__device__ __forceinline__ void cta_span(const RowSched& s, long long& u0, long long& u1) {
    span_range(s, blockIdx.x, u0, u1);
}

// the span that contains unit u (inverse of the floor partition above)
__device__ __forceinline__ int span_owner(const RowSched& s, long long u) {
    return (int)(((u + 1) * (long long)s.nspans - 1) / s.U);
}

// torch-style index: negative wraps once, then clamp for memory safety (eager would raise).
__device__ __forceinline__ int wrap_timestep(long long t, int T) {
    if (t < 0) t += T;
    if (t < 0) t = 0;
    if (t >= T) t = T - 1;
    return (int)t;
}

// Workspace layout — every position is INDEPENDENT of the per-call B, so one zero-initialised
// workspace can be reused across calls with different batch sizes (the kernels leave it clean):
//   [0, kSlotBytes)            partial slots, 2 per span: a span has at most two row segments that
//                              share their row with another span (its first and its last), all rows
//                              in between are complete inside the span and need no slot.
//                                slot(span c, segment) = 2c      if the row began before the span
//                                                        2c + 1  if the row begins inside it
//   [kSlotBytes, +4)           span-claim counter (dynamic scheduling)
//   [kSlotBytes + 4*(1+r))     ticket counter of row r
struct RowWorkspace {
    unsigned int* counters;   // [0] span-claim counter, [1 + r] ticket of row r
    float* partials;
};

constexpr long long kSlotBytes = 2LL * kMaxSpans * kRowPartialStride * (long long)sizeof(float);

inline long long row_ws_bytes(long long B) {
    return kSlotBytes + (((B + 1) * (long long)sizeof(unsigned int) + 255) / 256) * 256;
}

inline RowWorkspace carve_row_workspace(void* ws, long long /*B*/) {
    RowWorkspace r;
    r.partials = reinterpret_cast<float*>(ws);
    r.counters = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(ws) + kSlotBytes);
    return r;
}

__device__ __forceinline__ long long partial_slot(long long span, bool row_began_before_span) {
    return 2 * span + (row_began_before_span ? 0 : 1);
}

// 128-bit L2-coherent load/store of a partial slot (other SMs wrote it: bypass L1)
__device__ __forceinline__ float4 ld_slot(const float* p) {
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ void st_slot(float* p, float a, float b, float c) {
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(a), "f"(b), "f"(c), "f"(0.f) : "memory");
}

// Reduce K (<= 3) per-thread sums over the CTA and, if the row is shared with other CTAs, over
// all contributors. Returns true in exactly the threads of warp 0 of the CTA that ends up holding
// the complete row totals (`tot`): the only contributor, or the last one to arrive.
// Call from ALL threads (contains __syncthreads).
template <int K>
__device__ __forceinline__ bool row_reduce(float (&acc)[K], double (&tot)[K], const RowSched& s, const RowWorkspace& ws,
                                           long long row, bool row_began_before_span, float* red, int* flag) {
    static_assert(K <= 3, "slot holds 3 values");
    block_sum<K>(acc, red);
    const long long rs = row * s.upr;
    const int first = span_owner(s, rs), last = span_owner(s, rs + s.upr - 1);
    if (first == last) {
#pragma unroll
        for (int k = 0; k < K; ++k) tot[k] = (double)acc[k];
        return threadIdx.x < 32;
    }
    // Only thread 0 publishes, so only thread 0 needs the (expensive, store-draining) device fence.
    if (threadIdx.x == 0) {
        st_slot(ws.partials + partial_slot(blockIdx.x, row_began_before_span) * kRowPartialStride, acc[0],
                K > 1 ? acc[1] : 0.f, K > 2 ? acc[2] : 0.f);
        __threadfence();
        unsigned int* counter = ws.counters + 1 + row;
        const unsigned int tk = atomicAdd(counter, 1u);
        const int is_last = (tk == (unsigned)(last - first));
        if (is_last) *counter = 0u;  // leave the workspace clean for the next launch
        *flag = is_last;
    }
    __syncthreads();
    if (*flag == 0 || threadIdx.x >= 32) return false;
    __threadfence();  // acquire side
    double t[3] = {0.0, 0.0, 0.0};
    const int n = last - first + 1;
    for (int i = threadIdx.x; i < n; i += 32) {   // contributors are the consecutive spans first..last
        const float4 v = ld_slot(ws.partials + partial_slot(first + i, i != 0) * kRowPartialStride);
        t[0] += (double)v.x; t[1] += (double)v.y; t[2] += (double)v.z;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) tot[k] = warp_sum(t[k]);
    return true;
}

template <typename T, int W>
__device__ __forceinline__ void store_unit(T* p, const float (&f)[W]) {
    if constexpr (W == 1) {
        VecTraits<T>::store1(p, f[0]);
    } else {
        static_assert(W == VecTraits<T>::N, "vector width");
        stg_stream(p, VecTraits<T>::pack(f));
    }
}

// Raw fetch used to batch all loads of an iteration before the first use (MLP).
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

// Row segment of the CTA span [u0, u1) that lies in `row`; unit indices relative to the row start.
struct RowSeg { long long begin, end; };
__device__ __forceinline__ RowSeg row_segment(const RowSched& s, long long u0, long long u1, long long row) {
    const long long rb = row * s.upr;
    RowSeg g;
    g.begin = (u0 > rb ? u0 - rb : 0);
    g.end = (u1 < rb + s.upr ? u1 - rb : s.upr);
    return g;
}

}  // namespace siss
