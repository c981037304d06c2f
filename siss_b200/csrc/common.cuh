// Shared device helpers for the SISS hot-path kernels (sm_100a only).
//
// Every kernel in this directory is HBM-bound streaming work (<= ~10 flop/byte), so the
// helpers here are about moving bytes: 128-bit coalesced global accesses that bypass L1,
// dtype pack/unpack at the register level, and fixed-order (deterministic) reductions.
// No tensor cores on this path by design (BASELINE.json north_star).
#pragma once

#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "../../include/siss_b200.h"

namespace siss {

constexpr int kThreads = 256;          // CTA size for all streaming kernels
constexpr int kWarps = kThreads / 32;
constexpr int kNumSMsB200 = 148;

// ---------------------------------------------------------------------------------------
// 128-bit streaming global accesses. `.nc` + `L1::no_allocate`: every input byte on this
// path is touched exactly once per kernel, so L1 allocation is pure pollution.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// L2 eviction-priority policies for loads whose data a FOLLOWING kernel re-reads (K4a's tail, consumed first by K4b's
// reverse walk): `evict_last` lines are the last candidates for replacement, `evict_first` marks read-once data.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint4 ldg_stream_hint(const void* p, unsigned long long policy) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p), "l"(policy));
    return r;
}

// plain (coherent) 128-bit load: for buffers another kernel may still be L2-resident for
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---------------------------------------------------------------------------------------
// dtype traits: a 16-byte vector holds N = 16/sizeof(T) elements.
// All arithmetic is done in fp32 after unpacking; packing rounds to nearest even, which is
// what ATen's eager kernels do for bf16/fp16 outputs.
// ---------------------------------------------------------------------------------------
template <typename T> struct VecTraits;

template <> struct VecTraits<float> {
    static constexpr int N = 4;
    __device__ static __forceinline__ void unpack(const uint4& r, float (&f)[4]) {
        f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
        f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
    }
    __device__ static __forceinline__ uint4 pack(const float (&f)[4]) {
        return make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]),
                          __float_as_uint(f[2]), __float_as_uint(f[3]));
    }
    __device__ static __forceinline__ float round(float v) { return v; }
    // one unit of x_t = sa*x + s1*n in eager's order (mul, mul, add; no FMA contraction)
    __device__ static __forceinline__ uint4 noised_unit(float sa, float s1, const uint4& x, const uint4& n) {
        float xf[4], nf[4], o[4];
        unpack(x, xf); unpack(n, nf);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __fadd_rn(__fmul_rn(sa, xf[i]), __fmul_rn(s1, nf[i]));
        return pack(o);
    }
    __device__ static __forceinline__ float load1(const float* p) { return *p; }
    __device__ static __forceinline__ void store1(float* p, float v) { *p = v; }
};

template <> struct VecTraits<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void unpack(const uint4& r, float (&f)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i]     = __uint_as_float(w[i] << 16);          // low half  = element 2i
            f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);  // high half = element 2i+1
        }
    }
    __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ static __forceinline__ float round(float v) {
        return __bfloat162float(__float2bfloat16_rn(v));
    }
    // Packed native bf16 arithmetic (HMUL2/HADD2.BF16, two elements per instruction, no cvt): each
    // op rounds once to bf16, which equals eager's "compute in fp32, round to bf16" because the fp32
    // product of two bf16 values is exact and the fp32 sum cannot land on a bf16 tie unless the exact
    // sum does. The _rn intrinsics forbid contraction into HFMA2. sa/s1 are bf16-representable.
    __device__ static __forceinline__ uint4 noised_unit(float sa, float s1, const uint4& x, const uint4& n) {
        const __nv_bfloat162 sa2 = __float2bfloat162_rn(sa), s12 = __float2bfloat162_rn(s1);
        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, nw[4] = {n.x, n.y, n.z, n.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 xv = *reinterpret_cast<const __nv_bfloat162*>(&xw[i]);
            const __nv_bfloat162 nv = *reinterpret_cast<const __nv_bfloat162*>(&nw[i]);
            const __nv_bfloat162 r = __hadd2_rn(__hmul2_rn(sa2, xv), __hmul2_rn(s12, nv));
            o[i] = *reinterpret_cast<const uint32_t*>(&r);
        }
        return make_uint4(o[0], o[1], o[2], o[3]);
    }
    __device__ static __forceinline__ float load1(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static __forceinline__ void store1(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};

template <> struct VecTraits<__half> {
    static constexpr int N = 8;
    __device__ static __forceinline__ void unpack(const uint4& r, float (&f)[8]) {
        const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
            f[2 * i] = v.x; f[2 * i + 1] = v.y;
        }
    }
    __device__ static __forceinline__ uint4 pack(const float (&f)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __half2 h = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
    __device__ static __forceinline__ float round(float v) { return __half2float(__float2half_rn(v)); }
    // Packed native fp16 arithmetic; same argument as for bf16 (11-bit significands: exact fp32 product).
    __device__ static __forceinline__ uint4 noised_unit(float sa, float s1, const uint4& x, const uint4& n) {
        const __half2 sa2 = __float2half2_rn(sa), s12 = __float2half2_rn(s1);
        const uint32_t xw[4] = {x.x, x.y, x.z, x.w}, nw[4] = {n.x, n.y, n.z, n.w};
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __half2 xv = *reinterpret_cast<const __half2*>(&xw[i]);
            const __half2 nv = *reinterpret_cast<const __half2*>(&nw[i]);
            const __half2 r = __hadd2_rn(__hmul2_rn(sa2, xv), __hmul2_rn(s12, nv));
            o[i] = *reinterpret_cast<const uint32_t*>(&r);
        }
        return make_uint4(o[0], o[1], o[2], o[3]);
    }
    __device__ static __forceinline__ float load1(const __half* p) { return __half2float(*p); }
    __device__ static __forceinline__ void store1(__half* p, float v) { *p = __float2half_rn(v); }
};

// ---------------------------------------------------------------------------------------
// Reductions. Warp level: xor-butterfly shuffles (all lanes end with the total).
// CTA level: one smem slot per warp, summed in warp order by every thread (fixed order, so
// bitwise reproducible run to run).
// ---------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Reduce K values per thread across the CTA. `smem` must hold K * kWarps elements of T.
// Returns with every thread holding the K totals. Contains two __syncthreads().
template <int K, typename T>
__device__ __forceinline__ void block_sum(T (&v)[K], T* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();  // protect smem from a previous use
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * kWarps + warp] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        T acc = smem[k * kWarps];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) acc += smem[k * kWarps + w];
        v[k] = acc;
    }
}

// "last CTA done" ticket: returns true in exactly one CTA per group, after every other CTA of
// the group has published (release) its partials. Call from ALL threads of the CTA; the
// result is CTA-uniform. The winner resets the counter so the workspace is reusable without
// a memset between launches.
__device__ __forceinline__ bool last_cta_ticket(unsigned int* counter, unsigned int group_size, int* smem_flag) {
    __threadfence();            // publish this CTA's partial stores device-wide
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        int last = (t == group_size - 1u);
        if (last) *counter = 0u;  // nobody else touches it until the next launch
        *smem_flag = last;
    }
    __syncthreads();
    const bool last = (*smem_flag != 0);
    if (last) __threadfence();  // acquire side: see the other CTAs' partials
    return last;
}

inline int round_up_div(long long a, long long b) { return (int)((a + b - 1) / b); }

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace siss
