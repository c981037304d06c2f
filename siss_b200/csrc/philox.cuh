// Counter-based device RNG (opt-in): Philox4x32-10 + Box-Muller.
// Stream definition — MUST match oracle/philox.py (the CPU restatement the tests compare against):
//   key     = (seed lo, seed hi)
//   noise   : counter = (c lo, c hi, draw lo, draw hi [bit 31 clear]), c = global element index / 4; the four
//             words of call c give elements 4c..4c+3: (z0,z1) = box_muller(w0,w1), (z2,z3) = box_muller(w2,w3)
//   per row : counter = (row lo, row hi, draw lo, draw hi | 0x80000000)
//   aux     : counter = (c lo, c hi, draw lo, draw hi | 0x40000000) — a second element-indexed tensor of the same draw
//             (EraseDiff's uniform forget target): element e takes word e % 4 of call e / 4, u = (w >> 8) * 2^-24 in [0, 1)
// draw < 2^62 (bits 30 and 31 of the high word select the domain).
// Every value is a pure function of (seed, draw, global index): independent of grid, scheduling, sharding.
#pragma once

#include "common.cuh"

namespace siss {

struct RngStream {
    uint32_t k0, k1;   // key   = seed
    uint32_t d0, d1;   // draw  (d1 bits 31 / 30 select the per-row / aux domain)
};

enum RngDomain : uint32_t { kRngNoise = 0u, kRngRows = 0x80000000u, kRngAux = 0x40000000u };

__host__ __device__ inline RngStream make_rng_stream(uint64_t seed, uint64_t draw, uint32_t domain) {
    RngStream r;
    r.k0 = (uint32_t)seed; r.k1 = (uint32_t)(seed >> 32);
    r.d0 = (uint32_t)draw; r.d1 = ((uint32_t)(draw >> 32) & 0x3FFFFFFFu) | domain;
    return r;
}

// `draw` held in DEVICE memory (advanced on the stream with siss_counter_add): a captured CUDA graph then draws fresh
// values on every replay, like the optimiser's device-side step counter.
__device__ __forceinline__ void rng_draw_from_device(RngStream& s, const unsigned long long* d_draw) {
    if (d_draw != nullptr) {
        const unsigned long long d = *d_draw;
        s.d0 = (uint32_t)d;
        s.d1 = ((uint32_t)(d >> 32) & 0x3FFFFFFFu) | (s.d1 & 0xC0000000u);
    }
}

__device__ __forceinline__ void philox4x32_10(const RngStream& s, unsigned long long index, uint32_t (&w)[4]) {
    uint32_t c0 = (uint32_t)index, c1 = (uint32_t)(index >> 32), c2 = s.d0, c3 = s.d1;
    uint32_t k0 = s.k0, k1 = s.k1;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        if (r) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1;
        c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    }
    w[0] = c0; w[1] = c1; w[2] = c2; w[3] = c3;
}

// (w + 0.5) / 2^32 in (0, 1]: the product is exact (power-of-two scale), one rounding in the add
__device__ __forceinline__ float rng_uniform(uint32_t w) {
    return __fadd_rn(__fmul_rn(__uint2float_rn(w), 2.3283064365386963e-10f), 1.1641532182693481e-10f);
}

// 24-bit uniform in [0, 1), exactly representable (what torch.rand's fp32 values look like)
__device__ __forceinline__ float rng_uniform01(uint32_t w) { return __uint2float_rn(w >> 8) * 5.9604644775390625e-08f; }

// W consecutive uniforms in [0, 1) starting at global element e0 (aux domain stream).
template <int W>
__device__ __forceinline__ void rng_uniforms(const RngStream& s, unsigned long long e0, float (&u)[W]) {
    if constexpr (W == 1) {
        uint32_t w[4];
        philox4x32_10(s, e0 >> 2, w);
        const int lane = (int)(e0 & 3ull);
        u[0] = rng_uniform01(lane == 0 ? w[0] : lane == 1 ? w[1] : lane == 2 ? w[2] : w[3]);
    } else {
        static_assert(W % 4 == 0, "vector widths are multiples of one Philox call");
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            uint32_t w[4];
            philox4x32_10(s, (e0 >> 2) + i, w);
#pragma unroll
            for (int q = 0; q < 4; ++q) u[4 * i + q] = rng_uniform01(w[q]);
        }
    }
}

// Hardware transcendental units (MUFU.LG2 / MUFU.SIN / MUFU.COS): absolute error ~2^-21, far below what a
// noise sample needs, and ~6x fewer instructions than the IEEE-accurate library versions — the generator has
// to fit in the ALU headroom of a memory-bound kernel.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& z0, float& z1) {
    const float rad = sqrtf(-2.0f * __logf(rng_uniform(a)));
    float sn, cs;
    __sincosf(6.2831853071795865f * rng_uniform(b), &sn, &cs);
    z0 = rad * cs; z1 = rad * sn;
}

// W consecutive normals starting at global element e0. Fast path: e0 % 4 == 0 and W in {4, 8}.
template <int W>
__device__ __forceinline__ void rng_normals(const RngStream& s, unsigned long long e0, float (&z)[W]) {
    if constexpr (W == 1) {
        uint32_t w[4];
        philox4x32_10(s, e0 >> 2, w);
        const int lane = (int)(e0 & 3ull);
        float a, b;
        box_muller(lane < 2 ? w[0] : w[2], lane < 2 ? w[1] : w[3], a, b);
        z[0] = (lane & 1) ? b : a;
    } else {
        static_assert(W % 4 == 0, "vector widths are multiples of one Philox call");
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            uint32_t w[4];
            philox4x32_10(s, (e0 >> 2) + i, w);
            box_muller(w[0], w[1], z[4 * i + 0], z[4 * i + 1]);
            box_muller(w[2], w[3], z[4 * i + 2], z[4 * i + 3]);
        }
    }
}

}  // namespace siss
