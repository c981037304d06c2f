// Data-parallel gradient exchange through the NVSwitch's in-fabric reduction / replication (NVLS, multicast
// addresses; see p2p_common.cuh) fused with the two-term combine — and a PIPELINED form of the whole exchange.
//
// Bytes over each GPU's links, per parameter, N ranks (fp32):
//                                   outbound                         inbound
//   peer loads  (p2p reduce x2)     8 (N-1)/N  (peers pull from us)   8 (N-1)/N
//   peer stores (p2p gather)        4 (N-1)/N                         4 (N-1)/N      total per direction 12 (N-1)/N
//   multimem.ld_reduce x2           8          (switch reads all N)   8 / N
//   multimem.st                     4 / N                             4              total 8 + 4/N out, 4 + 8/N in
// so a reduce phase built on the switch loads the OUTBOUND links and leaves the inbound ones idle, and a gather phase
// does the opposite. The three-stage schedule (reduce both | scalars | combine + gather) cannot overlap them, because
// s and the clip need the global norms first. For the scaling-norm modes (SISS, SISS No-IS: s = scaling_norm / ||G_a||,
// delete_celeb.py:746) only ||G_a|| is needed before the combination can be FORMED; the clip is one scalar that can be
// applied afterwards. Pipelined schedule:
//
//   phase 1   reduce-scatter G_a (+ Saa)                         siss_p2p_reduce_norm3(x_mode = 2) or siss_nvls_reduce_norm3
//   barrier   (Saa of every rank visible)
//   phase 2   siss_nvls_xcombine_bcast: for the own shard, x = multimem.ld_reduce(G_x); y = x - s a (unclipped, the
//             reference's op order); accumulate Sxx, Sxa, Syy; multimem.st(y) back IN PLACE into every rank's G_x.
//             Reduce traffic (outbound) and gather traffic (inbound) run concurrently in ONE kernel:
//             4 + 4/N bytes per parameter in each direction.
//   barrier   (partial sums of every rank visible, every G_x complete)
//   phase 3   siss_scale_finalize: clip = min(1, max_norm / (sqrt(Syy) + 1e-6)) (clip_grad_norm_, delete_celeb.py:767),
//             stats, and G_x *= clip locally (8 B/param of HBM; skipped when clip == 1).
//
// In-place safety of phase 2: element e of every rank's G_x is read only by e's owner (through the multicast
// reduce) and overwritten only by the same thread afterwards (data dependence), so no rank can observe a mix.
// Element-wise the result is fl(fl(x - fl(s a)) clip), exactly the reference's three roundings; the norm of the
// combination is the directly accumulated fp64 sum of y^2 (what clip_grad_norm_ measures) rather than the algebraic
// Sxx - 2 s Sxa + s^2 Saa of the three-stage path: the two agree to fp64 rounding.

#include "p2p_common.cuh"

namespace siss {

// XMODE 0: x and a through multimem.ld_reduce; 1: x already reduced in shard_x (local), a through the switch;
// 2: a only (phase 1 of the pipelined exchange).
template <int U, int XMODE>
__global__ void __launch_bounds__(kThreads, kP2POcc)
nvls_reduce_norm3_kernel(const float* __restrict__ mc_x, const float* __restrict__ mc_a, int world, int rank,
                         long long shard_len, float* shard_x, float* __restrict__ shard_a,
                         double* __restrict__ sums3_local, PeerOut pub, P2PWorkspace ws) {
    __shared__ double red[3 * kWarps];
    __shared__ int flag;
    double acc[3] = {0.0, 0.0, 0.0};
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
            if (!ok[u]) continue;
            if (XMODE == 0) rx[u] = mc_ld_reduce_f32x4(mc_x + base_elem + 4 * i);
            else if (XMODE == 1) rx[u] = ldg_v4(shard_x + 4 * i);
            ra[u] = mc_ld_reduce_f32x4(mc_a + base_elem + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            if (XMODE == 0) stg_stream(shard_x + 4 * i, rx[u]);
            stg_stream(shard_a + 4 * i, ra[u]);
            float sx[4] = {0.f, 0.f, 0.f, 0.f}, sa[4];
            if (XMODE != 2) VecTraits<float>::unpack(rx[u], sx);
            VecTraits<float>::unpack(ra[u], sa);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const double ad = (double)sa[q];
                if (XMODE != 2) {
                    const double xd = (double)sx[q];
                    acc[0] = fma(xd, xd, acc[0]);
                    acc[2] = fma(xd, ad, acc[2]);
                }
                acc[1] = fma(ad, ad, acc[1]);
            }
        }
    }
    publish_rank_sums(acc, red, &flag, ws, sums3_local, pub, world, rank);
}

// Phase 2 of the pipelined exchange (scaling-norm modes). slots1: [world][4] doubles {_, Saa_r, _, _} of phase 1.
// Publishes {Sxx_r, Sxa_r, Syy_r} to slot [rank] of every peer's second slot array.
template <int U>
__global__ void __launch_bounds__(kThreads, kP2POcc)
nvls_xcombine_bcast_kernel(float* mc_x, const float* __restrict__ shard_a, const double* __restrict__ slots1,
                           int world, int rank, long long shard_len, float value, int inf_guard,
                           PeerOut pub2, P2PWorkspace ws) {
    __shared__ double red[3 * kWarps];
    __shared__ int flag;
    double saa = 0.0;
    for (int r = 0; r < world; ++r) saa += slots1[4 * r + 1];   // rank order: identical on every rank
    const float s = combine_scalars_from(0.0, saa, 0.0, SISS_COMBINE_SCALING_NORM, value, 0.0f, inf_guard, nullptr, false).s;
    double acc[3] = {0.0, 0.0, 0.0};   // Sxx, Sxa, Syy
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
            if (!ok[u]) continue;
            rx[u] = mc_ld_reduce_f32x4(mc_x + base_elem + 4 * i);
            ra[u] = ldg_stream(shard_a + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            float x[4], a[4], y[4];
            VecTraits<float>::unpack(rx[u], x);
            VecTraits<float>::unpack(ra[u], a);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                y[q] = __fsub_rn(x[q], __fmul_rn(s, a[q]));             // mul, sub — the clip's mul_ follows in phase 3
                const double xd = (double)x[q], ad = (double)a[q], yd = (double)y[q];
                acc[0] = fma(xd, xd, acc[0]);
                acc[1] = fma(xd, ad, acc[1]);
                acc[2] = fma(yd, yd, acc[2]);
            }
            mc_st_f32x4(mc_x + base_elem + 4 * i, VecTraits<float>::pack(y));                  // replicate into every rank's G_x, in place
        }
    }
    publish_rank_sums(acc, red, &flag, ws, nullptr, pub2, world, rank);
}

// Phase 3: scalars in the reference's fp32 op order from the two slot arrays, stats5, and the clip applied in place.
__global__ void __launch_bounds__(kThreads, 4)
scale_finalize_kernel(float* __restrict__ g, long long n, long long nvec, const double* __restrict__ slots1,
                      const double* __restrict__ slots2, int world, float value, float max_norm, int inf_guard,
                      float* __restrict__ stats5) {
    double saa = 0.0, sxx = 0.0, sxa = 0.0, syy = 0.0;
    for (int r = 0; r < world; ++r) {
        saa += slots1[4 * r + 1];
        sxx += slots2[4 * r + 0]; sxa += slots2[4 * r + 1]; syy += slots2[4 * r + 2];
    }
    (void)sxa;
    const float n_x = sqrtf((float)sxx), n_a = sqrtf((float)saa);
    const float s = combine_scalars_from(0.0, saa, 0.0, SISS_COMBINE_SCALING_NORM, value, 0.0f, inf_guard, nullptr, false).s;
    const float tn = (float)sqrt(syy);
    float clip = 1.0f;
    if (max_norm > 0.0f) {
        clip = __fdiv_rn(max_norm, __fadd_rn(tn, 1e-6f));   // torch.nn.utils.clip_grad_norm_
        clip = (clip > 1.0f) ? 1.0f : clip;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && stats5) {
        stats5[0] = n_x; stats5[1] = n_a; stats5[2] = s; stats5[3] = tn; stats5[4] = clip;
    }
    if (clip == 1.0f) return;   // x * 1.0f is the identity: clip_grad_norm_'s mul_ changes nothing
    constexpr int U = 4;
    const long long chunk = (long long)kThreads * U;
    const long long nchunks = (nvec + chunk - 1) / chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4 r[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            ok[u] = i < nvec;
            if (ok[u]) r[u] = ldg_v4(g + 4 * i);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = c * chunk + (long long)u * kThreads + threadIdx.x;
            float v[4];
            VecTraits<float>::unpack(r[u], v);
#pragma unroll
            for (int q = 0; q < 4; ++q) v[q] = __fmul_rn(v[q], clip);
            stg_stream(g + 4 * i, VecTraits<float>::pack(v));
        }
    }
    for (long long i = nvec * 4 + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
        g[i] = __fmul_rn(g[i], clip);
}

}  // namespace siss

using namespace siss;

extern "C" {

int siss_nvls_reduce_norm3(const float* mc_x, const float* mc_a, double* const* h_peer_scalars,
                           int world, int rank, int64_t shard_len, float* shard_x, float* shard_a,
                           double* sums3_local, int x_mode, void* workspace, siss_stream_t stream) {
    if (x_mode < 0 || x_mode > 2) return SISS_EINVAL;
    if ((x_mode == 0 && !mc_x) || !mc_a || !h_peer_scalars || (x_mode != 2 && !shard_x) || !shard_a || !sums3_local || !workspace)
        return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    if ((x_mode == 0 && !aligned16(mc_x)) || !aligned16(mc_a) || (x_mode != 2 && !aligned16(shard_x)) || !aligned16(shard_a))
        return SISS_EINVAL;
    PeerOut pub{};
    for (int r = 0; r < world; ++r) {
        if (!h_peer_scalars[r]) return SISS_EINVAL;
        pub.scalars[r] = h_peer_scalars[r];
    }
    P2PWorkspace ws = carve_p2p(workspace);
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int U = 4;
    const int grid = p2p_grid(shard_len / 4, U);
    if (x_mode == 0)
        nvls_reduce_norm3_kernel<U, 0><<<grid, kThreads, 0, st>>>(mc_x, mc_a, world, rank, shard_len, shard_x, shard_a, sums3_local, pub, ws);
    else if (x_mode == 1)
        nvls_reduce_norm3_kernel<U, 1><<<grid, kThreads, 0, st>>>(mc_x, mc_a, world, rank, shard_len, shard_x, shard_a, sums3_local, pub, ws);
    else
        nvls_reduce_norm3_kernel<U, 2><<<grid, kThreads, 0, st>>>(mc_x, mc_a, world, rank, shard_len, shard_x, shard_a, sums3_local, pub, ws);
    return (int)cudaGetLastError();
}

int siss_nvls_xcombine_bcast(float* mc_x, const float* shard_a, const double* scalar_slots1,
                             double* const* h_peer_scalars2, int world, int rank, int64_t shard_len,
                             float scaling_norm, int inf_guard, void* workspace, siss_stream_t stream) {
    if (!mc_x || !shard_a || !scalar_slots1 || !h_peer_scalars2 || !workspace) return SISS_EINVAL;
    if (world < 2 || world > kMaxWorld || rank < 0 || rank >= world || shard_len < 0 || shard_len % 4 != 0) return SISS_EINVAL;
    if (!aligned16(mc_x) || !aligned16(shard_a)) return SISS_EINVAL;
    PeerOut pub{};
    for (int r = 0; r < world; ++r) {
        if (!h_peer_scalars2[r]) return SISS_EINVAL;
        pub.scalars[r] = h_peer_scalars2[r];
    }
    constexpr int U = 4;
    nvls_xcombine_bcast_kernel<U><<<p2p_grid(shard_len / 4, U), kThreads, 0, (cudaStream_t)stream>>>(
        mc_x, shard_a, scalar_slots1, world, rank, shard_len, scaling_norm, inf_guard, pub, carve_p2p(workspace));
    return (int)cudaGetLastError();
}

int siss_scale_finalize(float* g, int64_t n, const double* scalar_slots1, const double* scalar_slots2, int world,
                        float scaling_norm, float max_norm, int inf_guard, float* stats5, siss_stream_t stream) {
    if (!g || n < 0 || !scalar_slots1 || !scalar_slots2 || world < 1 || world > kMaxWorld) return SISS_EINVAL;
    const long long nvec = aligned16(g) ? n / 4 : 0;
    long long work = (nvec + 1023) / 1024;
    long long grid = (long long)cached_sm_count() * 4;
    if (work < grid) grid = work;
    if (grid < 1) grid = 1;
    if (nvec == 0) {   // unaligned buffer: the scalar tail loop handles everything
        work = (n + kThreads - 1) / kThreads;
        grid = (long long)cached_sm_count() * 4;
        if (work < grid) grid = work;
        if (grid < 1) grid = 1;
    }
    scale_finalize_kernel<<<(int)grid, kThreads, 0, (cudaStream_t)stream>>>(g, n, nvec, scalar_slots1, scalar_slots2, world,
                                                                         scaling_norm, max_norm, inf_guard, stats5);
    return (int)cudaGetLastError();
}

}  // extern "C"
