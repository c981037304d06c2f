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

#include "p2p_common.cuh"
#include "adam.cuh"

namespace siss {

// U float4 units per thread per iteration; WORLD x 2 x U 128-bit loads in flight per thread.
// XMODE 0: pull G_x and G_a from every peer.
// XMODE 1: the G_x shard was already reduced (early reduce-scatter overlapped with the second backward): x is
//          read from the local shard and only G_a crosses NVLink here.
// XMODE 2: G_a only (first phase of the pipelined exchange, see nvls.cu): x is not touched, sums = {0, Saa, 0}.
template <int WORLD, int U, int XMODE>
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
        uint4 rx[XMODE == 0 ? WORLD : 1][U], ra[WORLD][U];
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
                if (XMODE == 0) rx[r][u] = ldg_v4(peers.x[r] + base_elem + 4 * i);
                else if (XMODE == 1 && r == 0) rx[0][u] = ldg_v4(shard_x + 4 * i);
                ra[r][u] = ldg_v4(peers.a[r] + base_elem + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            float sx[4] = {0.f, 0.f, 0.f, 0.f}, sa[4];
            if (XMODE != 2) VecTraits<float>::unpack(rx[0][u], sx);
            VecTraits<float>::unpack(ra[0][u], sa);
#pragma unroll
            for (int r = 1; r < WORLD; ++r) {   // fixed rank order
                float tx[4], ta[4];
                if (XMODE == 0) VecTraits<float>::unpack(rx[r][u], tx);
                VecTraits<float>::unpack(ra[r][u], ta);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (XMODE == 0) sx[q] = __fadd_rn(sx[q], tx[q]);
                    sa[q] = __fadd_rn(sa[q], ta[q]);
                }
            }
            if (XMODE == 0) stg_stream(shard_x + 4 * i, VecTraits<float>::pack(sx));
            stg_stream(shard_a + 4 * i, VecTraits<float>::pack(sa));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double xd = (double)sx[q], ad = (double)sa[q];
                if (XMODE != 2) { acc[0] = fma(xd, xd, acc[0]); acc[2] = fma(xd, ad, acc[2]); }
                acc[1] = fma(ad, ad, acc[1]);
            }
        }
    }
    publish_rank_sums(acc, red, &flag, ws, sums3_local, pub, WORLD, rank);
}

// MC: `peers.out[0]` is the MULTICAST address of the output buffer — one multimem.st replaces the WORLD peer stores.
template <int WORLD, int U, bool MC>
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
            if (MC) {
                mc_st_f32x4(peers.out[0] + base_elem + 4 * i, VecTraits<float>::pack(o));                              // all-gather in the switch
            } else {
                const uint4 v = VecTraits<float>::pack(o);
#pragma unroll
                for (int r = 0; r < WORLD; ++r) stg_stream(peers.out[r] + base_elem + 4 * i, v);  // all-gather by peer stores
            }
        }
    }
}

// Sharded (ZeRO-1) optimiser step fused with the parameter all-gather: K4b on this rank's gradient shard, the
// AdamW (+EMA) update of this rank's shard of parameters / moments in registers, and the NEW PARAMETERS stored
// to every peer's flat parameter buffer (instead of the combined gradient to every peer's G_x). Same outbound
// NVLink bytes as p2p_combine_allgather_kernel; the 40 B/param optimiser pass over the full buffer disappears
// from every rank (it is done once, on 1/WORLD of the parameters, here).
template <int WORLD, int U, bool EMA, bool MC>
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
            if (MC) {
                mc_st_f32x4(peers.out[0] + base_elem + 4 * i, VecTraits<float>::pack(p));                               // parameter all-gather (NVLS)
            } else {
                const uint4 pv = VecTraits<float>::pack(p);
#pragma unroll
                for (int r = 0; r < WORLD; ++r) stg_stream(peers.out[r] + base_elem + 4 * i, pv);  // parameter all-gather
            }
        }
    }
}

}  // namespace siss

using namespace siss;

extern "C" {

int64_t siss_p2p_workspace_bytes(void) { return 256 + (int64_t)kP2PMaxGrid * 3 * (int64_t)sizeof(double); }

int siss_p2p_reduce_norm3(const float* const* h_peers_x, const float* const* h_peers_a, double* const* h_peer_scalars,
                          int world, int rank, int64_t shard_len, float* shard_x, float* shard_a,
                          double* sums3_local, int x_prereduced, void* workspace, siss_stream_t stream) {
    if (x_prereduced < 0 || x_prereduced > 2) return SISS_EINVAL;
    if ((!x_prereduced && !h_peers_x) || !h_peers_a || !h_peer_scalars || (x_prereduced != 2 && !shard_x) || !shard_a ||
        !sums3_local || !workspace)
        return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    PeerPtrs peers{};
    PeerOut pub{};
    for (int r = 0; r < world; ++r) {
        if ((!x_prereduced && !h_peers_x[r]) || !h_peers_a[r] || !h_peer_scalars[r]) return SISS_EINVAL;
        if ((!x_prereduced && !aligned16(h_peers_x[r])) || !aligned16(h_peers_a[r])) return SISS_EINVAL;
        peers.x[r] = x_prereduced ? nullptr : h_peers_x[r]; peers.a[r] = h_peers_a[r]; pub.scalars[r] = h_peer_scalars[r];
    }
    if ((x_prereduced != 2 && !aligned16(shard_x)) || !aligned16(shard_a)) return SISS_EINVAL;
    P2PWorkspace ws = carve_p2p(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = shard_len / 4;
#define SISS_P2P_REDUCE(WORLD_, U_)                                                                                     \
    do {                                                                                                               \
        const int grid_ = p2p_grid(nvec, U_);                                                                          \
        if (x_prereduced == 1)                                                                                         \
            p2p_reduce_norm3_kernel<WORLD_, U_, 1><<<grid_, kThreads, 0, st>>>(peers, rank, shard_len, shard_x, shard_a, \
                                                                               sums3_local, pub, ws);                  \
        else if (x_prereduced == 2)                                                                                    \
            p2p_reduce_norm3_kernel<WORLD_, U_, 2><<<grid_, kThreads, 0, st>>>(peers, rank, shard_len, shard_x, shard_a, \
                                                                               sums3_local, pub, ws);                  \
        else                                                                                                           \
            p2p_reduce_norm3_kernel<WORLD_, U_, 0><<<grid_, kThreads, 0, st>>>(peers, rank, shard_len, shard_x, shard_a, \
                                                                               sums3_local, pub, ws);                  \
    } while (0)
    switch (world) {
        case 2: SISS_P2P_REDUCE(2, 4); break;
        case 4: SISS_P2P_REDUCE(4, 2); break;
        case 8: SISS_P2P_REDUCE(8, 1); break;
        default: return SISS_EUNSUPPORTED;  // 2, 4 or 8 GPUs of one NVSwitch box
    }
#undef SISS_P2P_REDUCE
    return (int)cudaGetLastError();
}

// Region-wise exchange (the reduce of each region of the flat buffer is issued as soon as autograd has finalised it):
// the per-region partial sums of this rank are added in region order and published to slot [rank] of every peer.
__global__ void publish_sums_kernel(const double* __restrict__ region_sums, int regions, PeerOut pub, int world, int rank) {
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int j = 0; j < regions; ++j) { t0 += region_sums[3 * j]; t1 += region_sums[3 * j + 1]; t2 += region_sums[3 * j + 2]; }
        for (int r = 0; r < world; ++r) {
            double* dst = pub.scalars[r] + 4 * rank;
            dst[0] = t0; dst[1] = t1; dst[2] = t2; dst[3] = 0.0;
        }
        __threadfence_system();
    }
}

static int launch_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                                    const PeerOut& peers, bool mc, int world, int rank, int64_t shard_len, int mode,
                                    float value, float max_norm, int inf_guard, float* stats5, cudaStream_t st) {
    const long long nvec = shard_len / 4;
#define SISS_CAG(WORLD_)                                                                                               \
    do {                                                                                                               \
        if (mc) p2p_combine_allgather_kernel<WORLD_, 4, true><<<p2p_grid(nvec, 4), kThreads, 0, st>>>(                 \
                    shard_x, shard_a, scalar_slots, rank, shard_len, peers, mode, value, max_norm, inf_guard, stats5); \
        else p2p_combine_allgather_kernel<WORLD_, 4, false><<<p2p_grid(nvec, 4), kThreads, 0, st>>>(                   \
                    shard_x, shard_a, scalar_slots, rank, shard_len, peers, mode, value, max_norm, inf_guard, stats5); \
    } while (0)
    switch (world) {
        case 2: SISS_CAG(2); break;
        case 4: SISS_CAG(4); break;
        case 8: SISS_CAG(8); break;
        default: return SISS_EUNSUPPORTED;
    }
#undef SISS_CAG
    return (int)cudaGetLastError();
}

int siss_publish_sums(const double* region_sums, int regions, double* const* h_peer_scalars, int world, int rank,
                      siss_stream_t stream) {
    if (!region_sums || regions < 1 || !h_peer_scalars || world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
        return SISS_EINVAL;
    PeerOut pub{};
    for (int r = 0; r < world; ++r) {
        if (!h_peer_scalars[r]) return SISS_EINVAL;
        pub.scalars[r] = h_peer_scalars[r];
    }
    publish_sums_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(region_sums, regions, pub, world, rank);
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
    return launch_combine_allgather(shard_x, shard_a, scalar_slots, peers, false, world, rank, shard_len, mode, value,
                                    max_norm, inf_guard, stats5, (cudaStream_t)stream);
}

int siss_nvls_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                                float* mc_out, int world, int rank, int64_t shard_len,
                                int mode, float value, float max_norm, int inf_guard, float* stats5,
                                siss_stream_t stream) {
    if (!shard_x || !shard_a || !scalar_slots || !mc_out || !aligned16(mc_out)) return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_NONE) return SISS_EINVAL;
    PeerOut peers{};
    peers.out[0] = mc_out;
    return launch_combine_allgather(shard_x, shard_a, scalar_slots, peers, true, world, rank, shard_len, mode, value,
                                    max_norm, inf_guard, stats5, (cudaStream_t)stream);
}

static int launch_adamw_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                                  const PeerOut& peers, bool mc, const float* p_local, int world, int rank,
                                  int64_t shard_len, int mode, float value, float max_norm, int inf_guard,
                                  float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                                  double weight_decay, int64_t step, const int64_t* d_step, const double* d_sched,
                                  float* ema_shard, double ema_decay, float* stats5, cudaStream_t st) {
    if (!shard_x || !shard_a || !scalar_slots || !exp_avg || !exp_avg_sq || !p_local) return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_ERASEDIFF || (step < 1 && !d_step)) return SISS_EINVAL;
    if (ema_shard && !d_sched && !(ema_decay >= 0.0 && ema_decay <= 1.0)) return SISS_EINVAL;
    if (!aligned16(shard_x) || !aligned16(shard_a) || !aligned16(exp_avg) || !aligned16(exp_avg_sq) ||
        !aligned16(ema_shard) || !aligned16(p_local))
        return SISS_EINVAL;
    long long hs;
    const AdamScalars as = make_adam_scalars(lr, beta1, beta2, eps, weight_decay, step, ema_decay, hs);
    const long long nvec = shard_len / 4;
    const long long* dstep = (const long long*)d_step;
#define SISS_LAUNCH_P2P_ADAMW(WORLD, EM, MC_)                                                                          \
    p2p_adamw_allgather_kernel<WORLD, 2, EM, MC_><<<p2p_grid(nvec, 2), kThreads, 0, st>>>(                              \
        shard_x, shard_a, scalar_slots, rank, shard_len, peers, p_local, exp_avg, exp_avg_sq, ema_shard, as, hs, dstep, \
        d_sched, mode, value, max_norm, inf_guard, stats5)
#define SISS_ADAMW_WORLD(WORLD)                                                                                        \
    do {                                                                                                               \
        if (ema_shard) { if (mc) SISS_LAUNCH_P2P_ADAMW(WORLD, true, true); else SISS_LAUNCH_P2P_ADAMW(WORLD, true, false); }   \
        else           { if (mc) SISS_LAUNCH_P2P_ADAMW(WORLD, false, true); else SISS_LAUNCH_P2P_ADAMW(WORLD, false, false); } \
    } while (0)
    switch (world) {
        case 2: SISS_ADAMW_WORLD(2); break;
        case 4: SISS_ADAMW_WORLD(4); break;
        case 8: SISS_ADAMW_WORLD(8); break;
        default: return SISS_EUNSUPPORTED;
    }
#undef SISS_ADAMW_WORLD
#undef SISS_LAUNCH_P2P_ADAMW
    return (int)cudaGetLastError();
}

int siss_p2p_adamw_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                             float* const* h_peers_param, int world, int rank, int64_t shard_len,
                             int mode, float value, float max_norm, int inf_guard,
                             float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int64_t step, const int64_t* d_step, const double* d_sched,
                             float* ema_shard, double ema_decay, float* stats5, siss_stream_t stream) {
    if (!h_peers_param || world < 2 || world > kMaxWorld || rank < 0 || rank >= world) return SISS_EINVAL;
    PeerOut peers{};
    for (int r = 0; r < world; ++r) {
        if (!h_peers_param[r] || !aligned16(h_peers_param[r])) return SISS_EINVAL;
        peers.out[r] = h_peers_param[r];
    }
    return launch_adamw_allgather(shard_x, shard_a, scalar_slots, peers, false,
                                  h_peers_param[rank] + (long long)rank * shard_len, world, rank, shard_len, mode, value,
                                  max_norm, inf_guard, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step,
                                  d_step, d_sched, ema_shard, ema_decay, stats5, (cudaStream_t)stream);
}

int siss_nvls_adamw_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                              float* mc_param, const float* param_local, int world, int rank, int64_t shard_len,
                              int mode, float value, float max_norm, int inf_guard,
                              float* exp_avg, float* exp_avg_sq, double lr, double beta1, double beta2, double eps,
                              double weight_decay, int64_t step, const int64_t* d_step, const double* d_sched,
                              float* ema_shard, double ema_decay, float* stats5, siss_stream_t stream) {
    if (!mc_param || !aligned16(mc_param) || !param_local || world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
        return SISS_EINVAL;
    PeerOut peers{};
    peers.out[0] = mc_param;
    return launch_adamw_allgather(shard_x, shard_a, scalar_slots, peers, true, param_local + (long long)rank * shard_len,
                                  world, rank, shard_len, mode, value, max_norm, inf_guard, exp_avg, exp_avg_sq, lr, beta1,
                                  beta2, eps, weight_decay, step, d_step, d_sched, ema_shard, ema_decay, stats5,
                                  (cudaStream_t)stream);
}

}  // extern "C"
