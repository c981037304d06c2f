// Fused compute + collective kernels over NVLink peer memory for the data-parallel gradient combine.
//
// The NCCL formulation of the exchange step is reduce-scatter(G_x), reduce-scatter(G_a), K4a,
// all-reduce(3 scalars), K4b, all-gather: five collectives around two kernels. Here each half is ONE
// kernel that moves its bytes over NVLink while it computes:
//
//   siss_p2p_reduce_norm3     rank r pulls elements [r*S, (r+1)*S) of G_x and G_a from EVERY rank's
//                             buffer (peer loads through the NVSwitch fabric, 128-bit, many in flight),
//                             sums them in rank order 0..N-1 (fixed order: deterministic, and every
//                             element is reduced by exactly one rank), keeps the reduced shard locally,
//                             accumulates the three fp64 sums of K4a on the fly, and its last CTA stores
//                             the rank's partial sums into every peer's scalar slot [r].
//   siss_p2p_combine_allgather  sums the N scalar slots in rank order (so all ranks derive bit-identical
//                             s and clip), evaluates K4b on the local reduced shard and stores each
//                             result vector into all N peers' output buffers (peer stores).
//
// Synchronisation between ranks (all buffers complete before a peer touches them) is the symmetric-
// memory barrier the host issues on the stream before, between and after the two kernels
// (siss_b200/grad_combine.py); inside the kernels everything is plain global loads/stores on mapped
// peer pointers. NVLink-bound: per GPU (N-1)/N * 8 B/param inbound for the first kernel,
// (N-1)/N * 4 B/param outbound for the second.

#include "common.cuh"
#include "combine_scalars.cuh"
#include "adam.cuh"

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

// U float4 units per thread per iteration; WORLD x 2 x U 128-bit loads in flight per thread.
// X_PRE: the G_x shard was already reduced (early reduce-scatter overlapped with the second backward): x is
// read from the local shard and only G_a crosses NVLink here.
template <int WORLD, int U, bool X_PRE>
__global__ void __launch_bounds__(kThreads, kP2POcc)
p2p_reduce_norm3_kernel(PeerPtrs peers, int rank, long long shard_len /* elements, % 4 == 0 */,
                        float* shard_x, float* __restrict__ shard_a,
                        double* __restrict__ sums3_local, PeerOut pub, P2PWorkspace ws) {
    __shared__ double red[3 * kWarps];
    __shared__ int flag;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long nvec = shard_len / 4;
    const long long base_elem = (long long)rank * shard_len;
    const long long chunk = (long long)kThreads * U;
    const long long nchunks = (nvec + chunk - 1) / chunk;

    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4 rx[WORLD][U], ra[WORLD][U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            ok[u] = i < nvec;
        }
        // all loads first: local rank's buffer comes from HBM, the others over NVLink
#pragma unroll
        for (int r = 0; r < WORLD; ++r) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
                if (!X_PRE) rx[r][u] = ldg_v4(peers.x[r] + base_elem + 4 * i);
                else if (r == 0) rx[0][u] = ldg_v4(shard_x + 4 * i);
                ra[r][u] = ldg_v4(peers.a[r] + base_elem + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            float sx[4], sa[4];
            VecTraits<float>::unpack(rx[0][u], sx);
            VecTraits<float>::unpack(ra[0][u], sa);
#pragma unroll
            for (int r = 1; r < WORLD; ++r) {   // fixed rank order
                float tx[4], ta[4];
                if (!X_PRE) VecTraits<float>::unpack(rx[r][u], tx);
                VecTraits<float>::unpack(ra[r][u], ta);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (!X_PRE) sx[q] = __fadd_rn(sx[q], tx[q]);
                    sa[q] = __fadd_rn(sa[q], ta[q]);
                }
            }
            if (!X_PRE) stg_stream(shard_x + 4 * i, VecTraits<float>::pack(sx));
            stg_stream(shard_a + 4 * i, VecTraits<float>::pack(sa));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double xd = (double)sx[q], ad = (double)sa[q];
                acc[0] = fma(xd, xd, acc[0]);
                acc[1] = fma(ad, ad, acc[1]);
                acc[2] = fma(xd, ad, acc[2]);
            }
        }
    }

    block_sum<3>(acc, red);
    if (threadIdx.x == 0) {
        ws.partials[3 * blockIdx.x + 0] = acc[0];
        ws.partials[3 * blockIdx.x + 1] = acc[1];
        ws.partials[3 * blockIdx.x + 2] = acc[2];
    }
    if (last_cta_ticket(ws.counter, gridDim.x, &flag)) {
        if (threadIdx.x < 32) {
            double t[3] = {0.0, 0.0, 0.0};
            const volatile double* p = ws.partials;
            for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) {
                t[0] += p[3 * b + 0]; t[1] += p[3 * b + 1]; t[2] += p[3 * b + 2];
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) t[k] = warp_sum(t[k]);
            if (threadIdx.x == 0) { sums3_local[0] = t[0]; sums3_local[1] = t[1]; sums3_local[2] = t[2]; }
            // publish this rank's partial sums into slot [rank] of every peer (and of itself)
            if (threadIdx.x < WORLD) {
                double* dst = pub.scalars[threadIdx.x] + 4 * rank;
                dst[0] = t[0]; dst[1] = t[1]; dst[2] = t[2]; dst[3] = 0.0;
                __threadfence_system();
            }
        }
    }
}

template <int WORLD, int U>
__global__ void __launch_bounds__(kThreads, kP2POcc)
p2p_combine_allgather_kernel(const float* __restrict__ shard_x, const float* __restrict__ shard_a,
                             const double* __restrict__ scalar_slots /* [WORLD][4] local */,
                             int rank, long long shard_len, PeerOut peers, int mode, float value, float max_norm,
                             int inf_guard, float* __restrict__ stats5) {
    double sxx = 0.0, saa = 0.0, sxa = 0.0;
#pragma unroll
    for (int r = 0; r < WORLD; ++r) {   // rank order: identical on every rank
        sxx += scalar_slots[4 * r + 0];
        saa += scalar_slots[4 * r + 1];
        sxa += scalar_slots[4 * r + 2];
    }
    const CombineScalars cs = combine_scalars_from(sxx, saa, sxa, mode, value, max_norm, inf_guard, stats5,
                                                   blockIdx.x == 0 && threadIdx.x == 0);
    const float s = cs.s, clip = cs.clip;
    const long long nvec = shard_len / 4;
    const long long base_elem = (long long)rank * shard_len;
    const long long chunk = (long long)kThreads * U;
    const long long nchunks = (nvec + chunk - 1) / chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4 rx[U], ra[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            ok[u] = i < nvec;
            if (ok[u]) { rx[u] = ldg_stream(shard_x + 4 * i); ra[u] = ldg_stream(shard_a + 4 * i); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            float x[4], a[4], o[4];
            VecTraits<float>::unpack(rx[u], x);
            VecTraits<float>::unpack(ra[u], a);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = __fmul_rn(__fsub_rn(x[q], __fmul_rn(s, a[q])), clip);
            const uint4 v = VecTraits<float>::pack(o);
#pragma unroll
            for (int r = 0; r < WORLD; ++r) stg_stream(peers.out[r] + base_elem + 4 * i, v);  // all-gather by peer stores
        }
    }
}

// Sharded (ZeRO-1) optimiser step fused with the parameter all-gather: K4b on this rank's gradient shard, the
// AdamW (+EMA) update of this rank's shard of parameters / moments in registers, and the NEW PARAMETERS stored
// to every peer's flat parameter buffer (instead of the combined gradient to every peer's G_x). Same outbound
// NVLink bytes as p2p_combine_allgather_kernel; the 40 B/param optimiser pass over the full buffer disappears
// from every rank (it is done once, on 1/WORLD of the parameters, here).
template <int WORLD, int U, bool EMA>
__global__ void __launch_bounds__(kThreads, kP2POcc)
p2p_adamw_allgather_kernel(const float* __restrict__ shard_x, const float* __restrict__ shard_a,
                           const double* __restrict__ scalar_slots, int rank, long long shard_len, PeerOut peers,
                           const float* __restrict__ p_local /* this rank's shard of its own parameter buffer */,
                           float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq, float* __restrict__ ema,
                           AdamScalars as, long long host_step, const long long* __restrict__ d_step,
                           const double* __restrict__ d_sched, int mode, float value, float max_norm, int inf_guard,
                           float* __restrict__ stats5) {
    if (d_sched != nullptr) adam_sched_from_device(as, d_sched, host_step);
    if (d_step != nullptr) adam_bias_from_step(as, *d_step);
    double sxx = 0.0, saa = 0.0, sxa = 0.0;
#pragma unroll
    for (int r = 0; r < WORLD; ++r) {   // rank order: identical on every rank
        sxx += scalar_slots[4 * r + 0];
        saa += scalar_slots[4 * r + 1];
        sxa += scalar_slots[4 * r + 2];
    }
    const CombineScalars cs = combine_scalars_from(sxx, saa, sxa, mode, value, max_norm, inf_guard, stats5,
                                                   blockIdx.x == 0 && threadIdx.x == 0);
    const float s = cs.s, clip = cs.clip;
    const long long nvec = shard_len / 4;
    const long long base_elem = (long long)rank * shard_len;
    const long long chunk = (long long)kThreads * U;
    const long long nchunks = (nvec + chunk - 1) / chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4 rx[U], ra[U], rp[U], rm[U], rv[U], re[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            ok[u] = i < nvec;
            if (ok[u]) {
                rx[u] = ldg_stream(shard_x + 4 * i); ra[u] = ldg_stream(shard_a + 4 * i);
                rp[u] = ldg_v4(p_local + 4 * i);
                rm[u] = ldg_v4(exp_avg + 4 * i); rv[u] = ldg_v4(exp_avg_sq + 4 * i);
                if (EMA) re[u] = ldg_v4(ema + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            float x[4], a[4], p[4], m[4], v[4];
            VecTraits<float>::unpack(rx[u], x);
            VecTraits<float>::unpack(ra[u], a);
            VecTraits<float>::unpack(rp[u], p);
            VecTraits<float>::unpack(rm[u], m);
            VecTraits<float>::unpack(rv[u], v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float g = __fmul_rn(__fsub_rn(x[q], __fmul_rn(s, a[q])), clip);
                adam_update(g, p[q], m[q], v[q], as);
            }
            stg_stream(exp_avg + 4 * i, VecTraits<float>::pack(m));
            stg_stream(exp_avg_sq + 4 * i, VecTraits<float>::pack(v));
            if (EMA) {
                float e[4];
                VecTraits<float>::unpack(re[u], e);
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = ema_update(e[q], p[q], as.ema_omd);
                stg_stream(ema + 4 * i, VecTraits<float>::pack(e));
            }
            const uint4 pv = VecTraits<float>::pack(p);
#pragma unroll
            for (int r = 0; r < WORLD; ++r) stg_stream(peers.out[r] + base_elem + 4 * i, pv);  // parameter all-gather
        }
    }
}

static int p2p_grid(long long nvec, int U) {
    const long long chunk = (long long)kThreads * U;
    long long work = (nvec + chunk - 1) / chunk;
    long long grid = (long long)cached_sm_count() * kP2POcc;
    if (grid > kP2PMaxGrid) grid = kP2PMaxGrid;
    if (work < grid) grid = work;
    if (grid < 1) grid = 1;
    return (int)grid;
}

}  // namespace siss

using namespace siss;

extern "C" {

int64_t siss_p2p_workspace_bytes(void) { return 256 + (int64_t)kP2PMaxGrid * 3 * (int64_t)sizeof(double); }

int siss_p2p_reduce_norm3(const float* const* h_peers_x, const float* const* h_peers_a, double* const* h_peer_scalars,
                          int world, int rank, int64_t shard_len, float* shard_x, float* shard_a,
                          double* sums3_local, int x_prereduced, void* workspace, siss_stream_t stream) {
    if ((!x_prereduced && !h_peers_x) || !h_peers_a || !h_peer_scalars || !shard_x || !shard_a || !sums3_local || !workspace)
        return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    PeerPtrs peers{};
    PeerOut pub{};
    for (int r = 0; r < world; ++r) {
        if ((!x_prereduced && !h_peers_x[r]) || !h_peers_a[r] || !h_peer_scalars[r]) return SISS_EINVAL;
        if ((!x_prereduced && !aligned16(h_peers_x[r])) || !aligned16(h_peers_a[r])) return SISS_EINVAL;
        peers.x[r] = x_prereduced ? nullptr : h_peers_x[r]; peers.a[r] = h_peers_a[r]; pub.scalars[r] = h_peer_scalars[r];
    }
    if (!aligned16(shard_x) || !aligned16(shard_a)) return SISS_EINVAL;
    P2PWorkspace ws = carve_p2p(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = shard_len / 4;
#define SISS_P2P_REDUCE(WORLD_, U_)                                                                                     \
    if (x_prereduced)                                                                                                  \
        p2p_reduce_norm3_kernel<WORLD_, U_, true><<<p2p_grid(nvec, U_), kThreads, 0, st>>>(peers, rank, shard_len, shard_x,  \
                                                                                            shard_a, sums3_local, pub, ws); \
    else                                                                                                               \
        p2p_reduce_norm3_kernel<WORLD_, U_, false><<<p2p_grid(nvec, U_), kThreads, 0, st>>>(peers, rank, shard_len, shard_x, \
                                                                                             shard_a, sums3_local, pub, ws)
    switch (world) {
        case 2: SISS_P2P_REDUCE(2, 4); break;
        case 4: SISS_P2P_REDUCE(4, 2); break;
        case 8: SISS_P2P_REDUCE(8, 1); break;
        default: return SISS_EUNSUPPORTED;  // 2, 4 or 8 GPUs of one NVSwitch box
    }
#undef SISS_P2P_REDUCE
    return (int)cudaGetLastError();
}

int siss_p2p_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                               float* const* h_peers_out, int world, int rank, int64_t shard_len,
                               int mode, float value, float max_norm, int inf_guard, float* stats5,
                               siss_stream_t stream) {
    if (!shard_x || !shard_a || !scalar_slots || !h_peers_out) return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_NONE) return SISS_EINVAL;
    PeerOut peers{};
    for (int r = 0; r < world; ++r) {
        if (!h_peers_out[r] || !aligned16(h_peers_out[r])) return SISS_EINVAL;
        peers.out[r] = h_peers_out[r];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = shard_len / 4;
    switch (world) {
        case 2: p2p_combine_allgather_kernel<2, 4><<<p2p_grid(nvec, 4), kThreads, 0, st>>>(shard_x, shard_a, scalar_slots, rank, shard_len, peers, mode, value, max_norm, inf_guard, stats5); break;
        case 4: p2p_combine_allgather_kernel<4, 4><<<p2p_grid(nvec, 4), kThreads, 0, st>>>(shard_x, shard_a, scalar_slots, rank, shard_len, peers, mode, value, max_norm, inf_guard, stats5); break;
        case 8: p2p_combine_allgather_kernel<8, 4><<<p2p_grid(nvec, 4), kThreads, 0, st>>>(shard_x, shard_a, scalar_slots, rank, shard_len, peers, mode, value, max_norm, inf_guard, stats5); break;
        default: return SISS_EUNSUPPORTED;
    }
    return (int)cudaGetLastError();
}

int siss_p2p_adamw_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                             float* const* h_peers_param, int world, int rank, int64_t shard_len,
                             int mode, float value, float max_norm, int inf_guard,
                             float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int64_t step, const int64_t* d_step, const double* d_sched,
                             float* ema_shard, double ema_decay, float* stats5, siss_stream_t stream) {
    if (!shard_x || !shard_a || !scalar_slots || !h_peers_param || !exp_avg || !exp_avg_sq) return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_ERASEDIFF || (step < 1 && !d_step)) return SISS_EINVAL;
    if (ema_shard && !d_sched && !(ema_decay >= 0.0 && ema_decay <= 1.0)) return SISS_EINVAL;
    if (!aligned16(shard_x) || !aligned16(shard_a) || !aligned16(exp_avg) || !aligned16(exp_avg_sq) || !aligned16(ema_shard))
        return SISS_EINVAL;
    PeerOut peers{};
    for (int r = 0; r < world; ++r) {
        if (!h_peers_param[r] || !aligned16(h_peers_param[r])) return SISS_EINVAL;
        peers.out[r] = h_peers_param[r];
    }
    long long hs;
    const AdamScalars as = make_adam_scalars(lr, beta1, beta2, eps, weight_decay, step, ema_decay, hs);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = shard_len / 4;
    const long long* dstep = (const long long*)d_step;
#define SISS_LAUNCH_P2P_ADAMW(WORLD, EM)                                                                               \
    p2p_adamw_allgather_kernel<WORLD, 2, EM><<<p2p_grid(nvec, 2), kThreads, 0, st>>>(                                   \
        shard_x, shard_a, scalar_slots, rank, shard_len, peers, h_peers_param[rank] + (long long)rank * shard_len,     \
        exp_avg, exp_avg_sq, ema_shard, as, hs, dstep, d_sched,                                                       \
        mode, value, max_norm, inf_guard, stats5)
    switch (world) {
        case 2: if (ema_shard) SISS_LAUNCH_P2P_ADAMW(2, true); else SISS_LAUNCH_P2P_ADAMW(2, false); break;
        case 4: if (ema_shard) SISS_LAUNCH_P2P_ADAMW(4, true); else SISS_LAUNCH_P2P_ADAMW(4, false); break;
        case 8: if (ema_shard) SISS_LAUNCH_P2P_ADAMW(8, true); else SISS_LAUNCH_P2P_ADAMW(8, false); break;
        default: return SISS_EUNSUPPORTED;
    }
#undef SISS_LAUNCH_P2P_ADAMW
    return (int)cudaGetLastError();
}

}  // extern "C"
