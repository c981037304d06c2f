// Shared by the peer-memory (p2p.cu) and multicast (nvls.cu) exchange kernels.
#pragma once

#include "common.cuh"
#include "combine_scalars.cuh"

namespace siss {

int cached_sm_count();

constexpr int kP2POcc = 2;
constexpr int kMaxWorld = 8;

struct P2PWorkspace {
    unsigned int* counter;
    double* partials;  // [grid][3]
};
constexpr int kP2PMaxGrid = 148 * 4;

inline P2PWorkspace carve_p2p(void* ws) {
    P2PWorkspace w;
    w.counter = reinterpret_cast<unsigned int*>(ws);
    w.partials = reinterpret_cast<double*>(reinterpret_cast<char*>(ws) + 256);
    return w;
}

struct PeerPtrs {
    const float* x[kMaxWorld];
    const float* a[kMaxWorld];
};
struct PeerOut {
    float* out[kMaxWorld];
    double* scalars[kMaxWorld];
};

inline int p2p_grid(long long nvec, int U) {
    const long long chunk = (long long)kThreads * U;
    long long work = (nvec + chunk - 1) / chunk;
    long long grid = (long long)cached_sm_count() * kP2POcc;
    if (grid > kP2PMaxGrid) grid = kP2PMaxGrid;
    if (work < grid) grid = work;
    if (grid < 1) grid = 1;
    return (int)grid;
}

// ---------------------------------------------------------------------------------------------------------
// NVLink SHARP (NVLS) through multicast addresses: one instruction addresses the same offset of EVERY rank's
// buffer. `multimem.ld_reduce` makes the NVSwitch fetch the 16 bytes from all ranks, add them in the switch and
// return ONE vector (inbound bytes / world); `multimem.st` sends one vector to the switch, which replicates it
// into every rank's buffer (outbound bytes / world). The order of the in-switch fp32 additions is fixed by the
// fabric, not by us: results are identical on every rank (each element is reduced once, by its owner) but are not
// bit-equal to the rank-ordered sum of the peer-load kernels.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 mc_ld_reduce_f32x4(const float* mc) {
    uint4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void mc_st_f32x4(float* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(mc), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// CTA epilogue of every reducing exchange kernel: block-reduce the three fp64 partial sums, publish them per CTA,
// and let the LAST CTA add the per-CTA partials in fixed order, store the rank's sums locally (`sums3_local`, may be
// null) and into slot [rank] (4 doubles) of every peer's scalar buffer — plain stores on mapped peer pointers,
// made visible by the symmetric-memory barrier that follows on the stream. Call from all threads.
__device__ __forceinline__ void publish_rank_sums(double (&acc)[3], double* red /* [3 * kWarps] */, int* flag,
                                                  const P2PWorkspace& ws, double* sums3_local, const PeerOut& pub,
                                                  int world, int rank) {
    block_sum<3>(acc, red);
    if (threadIdx.x == 0) {
        ws.partials[3 * blockIdx.x + 0] = acc[0];
        ws.partials[3 * blockIdx.x + 1] = acc[1];
        ws.partials[3 * blockIdx.x + 2] = acc[2];
    }
    if (last_cta_ticket(ws.counter, gridDim.x, flag)) {
        if (threadIdx.x < 32) {
            double t[3] = {0.0, 0.0, 0.0};
            const volatile double* p = ws.partials;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) {
                t[0] += p[3 * b + 0]; t[1] += p[3 * b + 1]; t[2] += p[3 * b + 2];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = warp_sum(t[k]);
            if (threadIdx.x == 0 && sums3_local) { sums3_local[0] = t[0]; sums3_local[1] = t[1]; sums3_local[2] = t[2]; }
            if ((int)threadIdx.x < world) {
                double* dst = pub.scalars[threadIdx.x] + 4 * rank;
                dst[0] = t[0]; dst[1] = t[1]; dst[2] = t[2]; dst[3] = 0.0;
                __threadfence_system();
            }
        }
    }
}

// Chunked variant for the copy-engine schedules (ce.cu), where one logical reduce is a SEQUENCE of launches on one
// stream (chunk c is reduced while chunk c+1 is still being copied): the last CTA of each launch adds the launch's
// total to the running totals `run3` (device memory; overwritten by the first chunk) in launch order — still a fixed
// summation order — and only the last chunk's launch publishes.
__device__ __forceinline__ void publish_rank_sums_chunked(double (&acc)[3], double* red, int* flag, const P2PWorkspace& ws,
                                                          double* run3, int first, int last, double* sums3_local,
                                                          const PeerOut& pub, int world, int rank) {
    block_sum<3>(acc, red);
    if (threadIdx.x == 0) {
        ws.partials[3 * blockIdx.x + 0] = acc[0];
        ws.partials[3 * blockIdx.x + 1] = acc[1];
        ws.partials[3 * blockIdx.x + 2] = acc[2];
    }
    if (last_cta_ticket(ws.counter, gridDim.x, flag)) {
        if (threadIdx.x < 32) {
            double t[3] = {0.0, 0.0, 0.0};
            const volatile double* p = ws.partials;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) {
                t[0] += p[3 * b + 0]; t[1] += p[3 * b + 1]; t[2] += p[3 * b + 2];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = warp_sum(t[k]);
            if (!first) {
                const volatile double* r = run3;
                t[0] += r[0]; t[1] += r[1]; t[2] += r[2];
            }
            __syncwarp();
            if (threadIdx.x == 0) {
                run3[0] = t[0]; run3[1] = t[1]; run3[2] = t[2];
                if (last && sums3_local) { sums3_local[0] = t[0]; sums3_local[1] = t[1]; sums3_local[2] = t[2]; }
            }
            if (last && (int)threadIdx.x < world) {
                double* dst = pub.scalars[threadIdx.x] + 4 * rank;
                dst[0] = t[0]; dst[1] = t[1]; dst[2] = t[2]; dst[3] = 0.0;
                __threadfence_system();
            }
        }
    }
}

}  // namespace siss
