// TMA bulk-copy -> shared-memory ring for the row-streaming kernels (sm_100a).
//
// Why: the row kernels do ~25-40 instructions per element, so with plain LDG every CTA alternates
// between "all loads in flight" and "all lanes computing", and at the batch sizes the reference
// uses (B = 4..64 per step) there are too few resident warps to hide that: ncu showed DRAM at
// 35-42 % with 25-34 % warps active. Here one producer thread per CTA issues 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx, SASS UBLKCP) kStages stages ahead of eight consumer
// warps, so the bytes in flight per SM are  CTAs/SM x (kStages-1) x stage bytes  (100+ KB)
// independent of register pressure, and HBM stays busy while the consumers compute.
//
//   full[s]  : count 1 (producer's arrive.expect_tx) + transaction bytes of the stage's copies
//   empty[s] : count kWarps (one elected lane per consumer warp arrives after its LDS are done)
//
// A stage holds kThreads units of every input stream, laid out exactly as in global memory, so
// consumer thread t reads unit t with one (16 B units) or two (32 B units) LDS.128.
#pragma once

#include "rowtile.cuh"

namespace siss {

constexpr int kPipeThreads = kThreads + 32;  // 8 consumer warps + 1 producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// 1-D bulk async copy global -> shared, completion counted in bytes on `bar`. `bytes` % 16 == 0,
// both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint4 lds128(const void* p) {
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(p)));
    return r;
}
// barrier over the 256 consumer threads only (the producer warp never joins it)
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kThreads) : "memory"); }

// Consumer-side CTA reduction with the named barrier (same fixed order as block_sum).
template <int K>
__device__ __forceinline__ void consumer_block_sum(float (&v)[K], float* smem) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = warp_sum(v[k]);
    consumer_sync();
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) smem[k * kWarps + warp] = v[k];
    }
    consumer_sync();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float acc = smem[k * kWarps];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) acc += smem[k * kWarps + w];
        v[k] = acc;
    }
}

// row_reduce for the consumer threads of a pipelined kernel (see rowtile.cuh::row_reduce).
template <int K>
__device__ __forceinline__ bool consumer_row_reduce(float (&acc)[K], double (&tot)[K], const RowSched& s,
                                                    const RowWorkspace& ws, long long row, bool row_began_before_span,
                                                    float* red, int* flag) {
    static_assert(K <= 3, "slot holds 3 values");
    consumer_block_sum<K>(acc, red);
    const long long rs = row * s.upr;
    const int first = span_owner(s, rs), last = span_owner(s, rs + s.upr - 1);
    if (first == last) {
#pragma unroll
        for (int k = 0; k < K; ++k) tot[k] = (double)acc[k];
        return threadIdx.x < 32;
    }
    if (threadIdx.x == 0) {
        st_slot(ws.partials + partial_slot(blockIdx.x, row_began_before_span) * kRowPartialStride, acc[0],
                K > 1 ? acc[1] : 0.f, K > 2 ? acc[2] : 0.f);
        __threadfence();
        unsigned int* counter = ws.counters + 1 + row;
        const unsigned int tk = atomicAdd(counter, 1u);
        const int is_last = (tk == (unsigned)(last - first));
        if (is_last) *counter = 0u;
        *flag = is_last;
    }
    consumer_sync();
    if (*flag == 0 || threadIdx.x >= 32) return false;
    __threadfence();
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

// ------------------------------------------------------------------------------------------------
// Generic pipelined row kernel.
//
// Op must provide:
//   static constexpr int NIN                 number of input streams (<= 4)
//   static constexpr int ub(int i)           bytes per unit of stream i (16 or 32)
//   static constexpr int K                   per-row sums (0..3)
//   static constexpr int kStages, kOcc
//   struct Params                            kernel arguments (POD)
//   struct Row                               per-row scalars
//   __device__ static Row  row_begin(const Params&, long long row)
//   __device__ static const char* stream(const Params&, long long row, int i)   base pointer of stream i
//   __device__ static void unit(const Params&, const Row&, const uint4 (&in)[NIN][2], long long unit_index,
//                               float (&acc)[K ? K : 1])      compute + global stores for one unit
//   __device__ static void row_end(const Params&, const Row&, long long row, const double (&tot)[K ? K : 1])
// `unit_index` is the global unit number (row * upr + u); element offset = unit_index * W.
// ------------------------------------------------------------------------------------------------
template <class Op>
struct PipeSmem {
    __host__ __device__ static constexpr int stage_bytes() {
        int b = 0;
        for (int i = 0; i < Op::NIN; ++i) b += Op::ub(i) * kThreads;
        return b;
    }
    __host__ __device__ static constexpr int stream_offset(int i) {
        int b = 0;
        for (int j = 0; j < i; ++j) b += Op::ub(j) * kThreads;
        return b;
    }
    __host__ __device__ static constexpr int bytes() { return Op::kStages * stage_bytes() + 2 * Op::kStages * 8 + 64; }
};

template <class Op>
__global__ void __launch_bounds__(kPipeThreads, Op::kOcc)
pipe_row_kernel(typename Op::Params p, RowWorkspace ws, RowSched s) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int S = Op::kStages;
    constexpr int SB = PipeSmem<Op>::stage_bytes();
    unsigned char* stage_mem = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + S * SB);
    uint64_t* empty = full + S;
    __shared__ float red[3 * kWarps];
    __shared__ int flag;

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, s.release_all ? kThreads : kWarps); }
        mbar_fence_init();
    }
    __syncthreads();

    long long u0, u1;
    cta_span(s, u0, u1);
    if (u0 >= u1) return;
    const bool is_producer = threadIdx.x >= kThreads;

    if (is_producer) {
        if (threadIdx.x != kThreads) return;  // one elected thread issues all copies
        int stage = 0;
        uint32_t phase = 0;
        for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
            const RowSeg seg = row_segment(s, u0, u1, row);
            // stream base pointers only (no per-row scalar chain here: the first copies must go out
            // immediately; the consumers fetch their row scalars while the data is in flight)
            const char* src[Op::NIN];
#pragma unroll
            for (int i = 0; i < Op::NIN; ++i) src[i] = Op::stream(p, row, i) + (row * s.upr) * (long long)Op::ub(i);
            for (long long ub = seg.begin; ub < seg.end; ub += kThreads) {
                const uint32_t n = (uint32_t)((seg.end - ub) < kThreads ? (seg.end - ub) : kThreads);
                mbar_wait(empty + stage, phase ^ 1u);
                uint32_t tx = 0;
#pragma unroll
                for (int i = 0; i < Op::NIN; ++i) tx += n * Op::ub(i);
                mbar_arrive_expect_tx(full + stage, tx);
#pragma unroll
                for (int i = 0; i < Op::NIN; ++i)
                    bulk_g2s(stage_mem + stage * SB + PipeSmem<Op>::stream_offset(i), src[i] + ub * (long long)Op::ub(i),
                             n * Op::ub(i), full + stage);
                if (++stage == S) { stage = 0; phase ^= 1u; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers (threads 0..255)
    int stage = 0;
    uint32_t phase = 0;
    const int lane = threadIdx.x & 31;
    for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
        const RowSeg seg = row_segment(s, u0, u1, row);
        const typename Op::Row r = Op::row_begin(p, row);
        float acc[Op::K ? Op::K : 1];
#pragma unroll
        for (int k = 0; k < (Op::K ? Op::K : 1); ++k) acc[k] = 0.f;
        for (long long ub = seg.begin; ub < seg.end; ub += kThreads) {
            const long long u = ub + threadIdx.x;
            const bool ok = u < seg.end;
            mbar_wait(full + stage, phase);
            uint4 in[Op::NIN][2];
            if (ok) {
#pragma unroll
                for (int i = 0; i < Op::NIN; ++i) {
                    const unsigned char* base = stage_mem + stage * SB + PipeSmem<Op>::stream_offset(i) +
                                                threadIdx.x * Op::ub(i);
                    in[i][0] = lds128(base);
                    if (Op::ub(i) == 32) in[i][1] = lds128(base + 16);
                }
            }
            if (ok) Op::unit(p, r, in, row * s.upr + u, acc);
            // Release the stage only AFTER the unit's math and global stores: those truly depend on the
            // registers the LDS above fill, so the shared-memory reads have completed by now. Releasing
            // right after *issuing* the LDS raced with the producer's next bulk copy (async proxy) under
            // MIO back-pressure: ~80 corrupted units per 1.5 M in K3 at the celeb shape, caught by
            // tests/test_fullsize_gpu.py. The ring is 5-6 stages deep, so holding a stage for one unit's
            // compute costs nothing measurable.
            if (s.release_all) {
                mbar_arrive(empty + stage);                      // every reader releases for itself
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);       // elected lane releases for its warp
            }
            if (++stage == S) { stage = 0; phase ^= 1u; }
        }
        if constexpr (Op::K > 0) {
            double tot[Op::K];
            if (consumer_row_reduce<Op::K>(acc, tot, s, ws, row, seg.begin > 0, red, &flag) && threadIdx.x == 0)
                Op::row_end(p, r, row, tot);
        }
    }
}

template <class Op>
static int launch_pipe(const typename Op::Params& p, RowWorkspace ws, long long B, long long D, int W, cudaStream_t st) {
    static bool configured = false;
    constexpr int smem = PipeSmem<Op>::bytes();
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pipe_row_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured = true;
    }
    RowSched s = make_row_sched(B, D, W, Op::kOcc);
    static const int release_all = env_int("SISS_RELEASE_ALL", 1);   // see the release note in pipe_row_kernel
    s.release_all = release_all == 1 ? 1 : 0;
    pipe_row_kernel<Op><<<s.grid, kPipeThreads, smem, st>>>(p, ws, s);
    return (int)cudaGetLastError();
}

}  // namespace siss
