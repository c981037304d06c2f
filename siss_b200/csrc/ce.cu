// Copy-engine schedules of the data-parallel gradient exchange.
//
// Measured on 2x B200 (profiles/r2_exchange_probe_w2.json, r2_ce_probe_w2.json): SM-issued peer traffic — the fused
// peer-load / peer-store kernels of p2p.cu and the multicast kernels of nvls.cu alike — tops out at 530-620 GB/s per
// direction, while a DMA copy of a mapped peer buffer (cudaMemcpyAsync, copy engines) moves 700-730 GB/s per direction
// with both directions busy. So here the BYTES go through the copy engines and the SMs only ever touch local HBM:
//
//   reduce      every rank DMA-pulls its shard of every peer's G_x / G_a into local staging, in `chunks` pieces on
//               one side stream per peer; as soon as piece c of all peers has landed, a kernel sums the WORLD
//               contributions of that piece in rank order (bit-identical to the peer-load kernel), writes the reduced
//               shard and accumulates the three fp64 sums — while piece c+1 is still on the wire.
//   gather      K4b on piece c of the shard writes into the rank's own G_x, and the piece is DMA-pushed to every peer
//               while K4b works on piece c+1.
//   pipelined   the G_a phase of the pipelined schedule of nvls.cu (x_mode 2) can use this reduce as well: DMA for the
//               phase whose traffic is symmetric anyway, the switch for the phase whose directions can be overlapped.
//               (A DMA form of that second phase would gain nothing: unicast pulls and pushes load BOTH directions
//               equally, so overlapping them moves the same bytes per direction as running them back to back.)
//
// Side streams and events belong to the library (one set per device, created on first use); everything is ordered
// against the caller's stream by events, nothing synchronises the host, and the fork/join pattern is CUDA-graph
// capturable. Rank-to-rank synchronisation (all inputs complete / all outputs landed) stays with the caller, exactly
// as for the siss_p2p_* pair.

#include "p2p_common.cuh"

namespace siss {

int current_device_slot();
constexpr int kMaxDevicesCe = 64;
constexpr int kMaxChunks = 16;

struct CeCtx {
    bool ready = false;
    cudaStream_t side[kMaxWorld];                // pulls from peer p
    cudaStream_t side2[kMaxWorld];               // second buffer's pulls from peer p (two DMA engines per peer)
    cudaStream_t push[kMaxWorld];                // pushes to peer p (separate: a push must not queue behind later pulls)
    cudaEvent_t fork;
    cudaEvent_t landed[kMaxWorld][kMaxChunks];   // piece c of peer p's data has arrived
    cudaEvent_t landed2[kMaxWorld][kMaxChunks];  // ... of the second buffer
    cudaEvent_t made[kMaxChunks];                // kernel of piece c has finished (gather side)
    cudaEvent_t joined[kMaxWorld];
};
static CeCtx g_ce[kMaxDevicesCe];

static int ce_ctx(CeCtx** out) {
    CeCtx& c = g_ce[current_device_slot() % kMaxDevicesCe];
    if (!c.ready) {
        cudaError_t e;
        for (int p = 0; p < kMaxWorld; ++p) {
            if ((e = cudaStreamCreateWithFlags(&c.side[p], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
            if ((e = cudaStreamCreateWithFlags(&c.push[p], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
            if ((e = cudaStreamCreateWithFlags(&c.side2[p], cudaStreamNonBlocking)) != cudaSuccess) return (int)e;
            if ((e = cudaEventCreateWithFlags(&c.joined[p], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
            for (int k = 0; k < kMaxChunks; ++k)
            {
                if ((e = cudaEventCreateWithFlags(&c.landed[p][k], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
                if ((e = cudaEventCreateWithFlags(&c.landed2[p][k], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
            }
        }
        for (int k = 0; k < kMaxChunks; ++k)
            if ((e = cudaEventCreateWithFlags(&c.made[k], cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        if ((e = cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming)) != cudaSuccess) return (int)e;
        c.ready = true;
    }
    *out = &c;
    return 0;
}

struct SrcPtrs {
    const float* x[kMaxWorld];   // start of rank r's contribution to THIS rank's shard (own buffer or staging)
    const float* a[kMaxWorld];
};

inline double* ce_run3(void* workspace) { return reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 128); }

// Sum the WORLD contributions of float4 units [vec_lo, vec_hi) of the shard in rank order (all local memory).
// XMODE as in p2p.cu: 0 = x and a, 1 = x already reduced in shard_x, 2 = a only.
template <int WORLD, int U, int XMODE>
__global__ void __launch_bounds__(kThreads, kP2POcc)
ce_reduce_chunk_kernel(SrcPtrs src, long long vec_lo, long long vec_hi, float* shard_x, float* __restrict__ shard_a,
                       double* run3, int first, int last, double* __restrict__ sums3_local, PeerOut pub, int rank,
                       P2PWorkspace ws) {
    __shared__ double red[3 * kWarps];
    __shared__ int flag;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long chunk = (long long)kThreads * U;
    const long long nchunks = (vec_hi - vec_lo + chunk - 1) / chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4 rx[XMODE == 0 ? WORLD : 1][U], ra[WORLD][U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) ok[u] = vec_lo + c * chunk + (long long)u * kThreads + threadIdx.x < vec_hi;
#pragma unroll
        for (int r = 0; r < WORLD; ++r) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                const long long i = vec_lo + c * chunk + (long long)u * kThreads + threadIdx.x;
                if (XMODE == 0) rx[r][u] = ldg_stream(src.x[r] + 4 * i);
                else if (XMODE == 1 && r == 0) rx[0][u] = ldg_v4(shard_x + 4 * i);
                ra[r][u] = ldg_stream(src.a[r] + 4 * i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = vec_lo + c * chunk + (long long)u * kThreads + threadIdx.x;
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
    publish_rank_sums_chunked(acc, red, &flag, ws, run3, first, last, sums3_local, pub, WORLD, rank);
}

// K4b on float4 units [vec_lo, vec_hi) of the shard, result into the rank's OWN output buffer (DMA-pushed afterwards).
template <int WORLD, int U>
__global__ void __launch_bounds__(kThreads, kP2POcc)
ce_combine_chunk_kernel(const float* __restrict__ shard_x, const float* __restrict__ shard_a,
                        const double* __restrict__ scalar_slots, long long vec_lo, long long vec_hi,
                        float* __restrict__ out_local /* own buffer + rank * shard_len */, int mode, float value,
                        float max_norm, int inf_guard, float* __restrict__ stats5, int write_stats) {
    double sxx = 0.0, saa = 0.0, sxa = 0.0;
#pragma unroll
    for (int r = 0; r < WORLD; ++r) {   // rank order: identical on every rank
        sxx += scalar_slots[4 * r + 0];
        saa += scalar_slots[4 * r + 1];
        sxa += scalar_slots[4 * r + 2];
    }
    const CombineScalars cs = combine_scalars_from(sxx, saa, sxa, mode, value, max_norm, inf_guard, stats5,
                                                   write_stats && blockIdx.x == 0 && threadIdx.x == 0);
    const float s = cs.s, clip = cs.clip;
    const long long chunk = (long long)kThreads * U;
    const long long nchunks = (vec_hi - vec_lo + chunk - 1) / chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        uint4 rx[U], ra[U];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = vec_lo + c * chunk + (long long)u * kThreads + threadIdx.x;
            ok[u] = i < vec_hi;
            if (ok[u]) { rx[u] = ldg_stream(shard_x + 4 * i); ra[u] = ldg_stream(shard_a + 4 * i); }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!ok[u]) continue;
            const long long i = vec_lo + c * chunk + (long long)u * kThreads + threadIdx.x;
            float x[4], a[4], o[4];
            VecTraits<float>::unpack(rx[u], x);
            VecTraits<float>::unpack(ra[u], a);
#pragma unroll
            for (int q = 0; q < 4; ++q) o[q] = __fmul_rn(__fsub_rn(x[q], __fmul_rn(s, a[q])), clip);
            stg_stream(out_local + 4 * i, VecTraits<float>::pack(o));
        }
    }
}

// Piece c of `chunks` with linearly graded sizes: `descending` (reduce side: chunks, chunks-1, ..., 1 parts) makes the
// LAST piece — whose kernel cannot overlap any copy — the smallest; ascending (gather side) makes the FIRST piece —
// whose kernel must finish before any push can start — the smallest.
static inline void piece(long long nvec, int chunks, int c, bool descending, long long& lo, long long& hi) {
    const long long total = (long long)chunks * (chunks + 1) / 2;
    auto prefix = [&](int k) -> long long {     // parts in pieces [0, k)
        return descending ? (long long)k * chunks - (long long)k * (k - 1) / 2 : (long long)k * (k + 1) / 2;
    };
    lo = nvec * prefix(c) / total;
    hi = nvec * prefix(c + 1) / total;
}

// DMA-pull pieces of this rank's shard from every peer into staging, one side stream per peer; `nbuf` buffers per piece.
static int ce_pull(CeCtx* cx, cudaStream_t st, int world, int rank, long long nvec, int chunks, long long base_elem,
                   const float* const* srcs0, float* staging0, const float* const* srcs1, float* staging1,
                   long long shard_len) {
    cudaError_t e;
    if ((e = cudaEventRecord(cx->fork, st)) != cudaSuccess) return (int)e;
    for (int k = 1; k < world; ++k) {
        const int p = (rank + k) % world;             // rotated: at any moment every rank pulls from a different peer
        const int j = p < rank ? p : p - 1;           // slot of peer p in the staging arrays
        if ((e = cudaStreamWaitEvent(cx->side[p], cx->fork, 0)) != cudaSuccess) return (int)e;
        if (srcs1 && (e = cudaStreamWaitEvent(cx->side2[p], cx->fork, 0)) != cudaSuccess) return (int)e;
        for (int c = 0; c < chunks; ++c) {
            long long lo, hi;
            piece(nvec, chunks, c, true, lo, hi);
            const size_t bytes = (size_t)(hi - lo) * 16;
            if (bytes && (e = cudaMemcpyAsync(staging0 + (long long)j * shard_len + 4 * lo, srcs0[p] + base_elem + 4 * lo,
                                              bytes, cudaMemcpyDeviceToDevice, cx->side[p])) != cudaSuccess) return (int)e;
            if ((e = cudaEventRecord(cx->landed[p][c], cx->side[p])) != cudaSuccess) return (int)e;
            if (srcs1) {
                if (bytes && (e = cudaMemcpyAsync(staging1 + (long long)j * shard_len + 4 * lo, srcs1[p] + base_elem + 4 * lo,
                                                  bytes, cudaMemcpyDeviceToDevice, cx->side2[p])) != cudaSuccess) return (int)e;
                if ((e = cudaEventRecord(cx->landed2[p][c], cx->side2[p])) != cudaSuccess) return (int)e;
            }
        }
    }
    return 0;
}

static int ce_wait_piece(CeCtx* cx, cudaStream_t st, int world, int rank, int c, bool two) {
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        cudaError_t e = cudaStreamWaitEvent(st, cx->landed[p][c], 0);
        if (e != cudaSuccess) return (int)e;
        if (two && (e = cudaStreamWaitEvent(st, cx->landed2[p][c], 0)) != cudaSuccess) return (int)e;
    }
    return 0;
}

// DMA-push piece c of the own output region to every peer after the kernel that produced it.
static int ce_push_piece(CeCtx* cx, cudaStream_t st, int world, int rank, long long nvec, int chunks, int c,
                         long long base_elem, float* const* outs) {
    cudaError_t e;
    if ((e = cudaEventRecord(cx->made[c], st)) != cudaSuccess) return (int)e;
    long long lo, hi;
    piece(nvec, chunks, c, false, lo, hi);
    const size_t bytes = (size_t)(hi - lo) * 16;
    for (int k = 1; k < world; ++k) {
        const int p = (rank + k) % world;
        if ((e = cudaStreamWaitEvent(cx->push[p], cx->made[c], 0)) != cudaSuccess) return (int)e;
        if (bytes && (e = cudaMemcpyAsync(outs[p] + base_elem + 4 * lo, outs[rank] + base_elem + 4 * lo, bytes,
                                          cudaMemcpyDeviceToDevice, cx->push[p])) != cudaSuccess) return (int)e;
    }
    return 0;
}

static int ce_join(CeCtx* cx, cudaStream_t st, int world, int rank) {
    cudaError_t e;
    for (int p = 0; p < world; ++p) {
        if (p == rank) continue;
        if ((e = cudaEventRecord(cx->joined[p], cx->push[p])) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(st, cx->joined[p], 0)) != cudaSuccess) return (int)e;
    }
    return 0;
}

static bool ce_args_ok(int world, int rank, int64_t shard_len, int chunks) {
    return (world == 2 || world == 4 || world == 8) && rank >= 0 && rank < world && shard_len >= 0 && shard_len % 4 == 0 &&
           chunks >= 1 && chunks <= kMaxChunks;
}

}  // namespace siss

using namespace siss;

extern "C" {

int siss_ce_reduce_norm3(const float* const* h_peers_x, const float* const* h_peers_a, double* const* h_peer_scalars,
                         int world, int rank, int64_t shard_len, float* staging, float* shard_x, float* shard_a,
                         double* sums3_local, int x_mode, int chunks, void* workspace, siss_stream_t stream) {
    if (x_mode < 0 || x_mode > 2 || !ce_args_ok(world, rank, shard_len, chunks)) return SISS_EINVAL;
    if ((x_mode == 0 && !h_peers_x) || !h_peers_a || !h_peer_scalars || !staging || (x_mode != 2 && !shard_x) || !shard_a ||
        !sums3_local || !workspace || !aligned16(staging) || !aligned16(shard_a) || (x_mode != 2 && !aligned16(shard_x)))
        return SISS_EINVAL;
    CeCtx* cx;
    int rc = ce_ctx(&cx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = shard_len / 4, base = (long long)rank * shard_len;
    float* stage_a = staging;
    float* stage_x = staging + (long long)(world - 1) * shard_len;
    if ((rc = ce_pull(cx, st, world, rank, nvec, chunks, base, h_peers_a, stage_a, x_mode == 0 ? h_peers_x : nullptr, stage_x,
                      shard_len)))
        return rc;
    SrcPtrs src{};
    PeerOut pub{};
    for (int r = 0; r < world; ++r) {
        if (!h_peers_a[r] || !h_peer_scalars[r] || (x_mode == 0 && !h_peers_x[r])) return SISS_EINVAL;
        const int j = r < rank ? r : r - 1;
        src.a[r] = (r == rank) ? h_peers_a[r] + base : stage_a + (long long)j * shard_len;
        src.x[r] = (x_mode != 0) ? nullptr : ((r == rank) ? h_peers_x[r] + base : stage_x + (long long)j * shard_len);
        pub.scalars[r] = h_peer_scalars[r];
    }
    P2PWorkspace ws = carve_p2p(workspace);
    double* run3 = ce_run3(workspace);
    for (int c = 0; c < chunks; ++c) {
        if ((rc = ce_wait_piece(cx, st, world, rank, c, x_mode == 0))) return rc;
        long long lo, hi;
        piece(nvec, chunks, c, true, lo, hi);
        const int first = c == 0, last = c == chunks - 1;
#define SISS_CE_REDUCE(WORLD_, U_)                                                                                      \
        do {                                                                                                           \
            const int grid_ = p2p_grid(hi - lo, U_);                                                                   \
            if (x_mode == 0) ce_reduce_chunk_kernel<WORLD_, U_, 0><<<grid_, kThreads, 0, st>>>(src, lo, hi, shard_x, shard_a, run3, first, last, sums3_local, pub, rank, ws); \
            else if (x_mode == 1) ce_reduce_chunk_kernel<WORLD_, U_, 1><<<grid_, kThreads, 0, st>>>(src, lo, hi, shard_x, shard_a, run3, first, last, sums3_local, pub, rank, ws); \
            else ce_reduce_chunk_kernel<WORLD_, U_, 2><<<grid_, kThreads, 0, st>>>(src, lo, hi, shard_x, shard_a, run3, first, last, sums3_local, pub, rank, ws); \
        } while (0)
        switch (world) {
            case 2: SISS_CE_REDUCE(2, 4); break;
            case 4: SISS_CE_REDUCE(4, 2); break;
            default: SISS_CE_REDUCE(8, 1); break;
        }
#undef SISS_CE_REDUCE
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}

int siss_ce_combine_allgather(const float* shard_x, const float* shard_a, const double* scalar_slots,
                              float* const* h_peers_out, int world, int rank, int64_t shard_len, int chunks,
                              int mode, float value, float max_norm, int inf_guard, float* stats5, siss_stream_t stream) {
    if (!shard_x || !shard_a || !scalar_slots || !h_peers_out || !ce_args_ok(world, rank, shard_len, chunks)) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_NONE) return SISS_EINVAL;
    for (int r = 0; r < world; ++r)
        if (!h_peers_out[r] || !aligned16(h_peers_out[r])) return SISS_EINVAL;
    CeCtx* cx;
    int rc = ce_ctx(&cx);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long nvec = shard_len / 4, base = (long long)rank * shard_len;
    float* out_local = h_peers_out[rank] + base;
    for (int c = 0; c < chunks; ++c) {
        long long lo, hi;
        piece(nvec, chunks, c, false, lo, hi);
        const int grid = p2p_grid(hi - lo, 4);
        switch (world) {
            case 2: ce_combine_chunk_kernel<2, 4><<<grid, kThreads, 0, st>>>(shard_x, shard_a, scalar_slots, lo, hi, out_local, mode, value, max_norm, inf_guard, stats5, c == 0); break;
            case 4: ce_combine_chunk_kernel<4, 4><<<grid, kThreads, 0, st>>>(shard_x, shard_a, scalar_slots, lo, hi, out_local, mode, value, max_norm, inf_guard, stats5, c == 0); break;
            default: ce_combine_chunk_kernel<8, 4><<<grid, kThreads, 0, st>>>(shard_x, shard_a, scalar_slots, lo, hi, out_local, mode, value, max_norm, inf_guard, stats5, c == 0); break;
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return (int)e;
        if ((rc = ce_push_piece(cx, st, world, rank, nvec, chunks, c, base, h_peers_out))) return rc;
    }
    return ce_join(cx, st, world, rank);
}

}  // extern "C"
