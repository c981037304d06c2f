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

constexpr int kPipeThreads = kThreads + 64;  // 8 consumer warps + 1 producer warp + 1 reducer warp

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

// ------------------------------------------------------------------------------------------------
// Row-segment reduction off the consumers' critical path.
//
// Per segment every consumer warp reduces its K sums with shuffles and lane 0 drops them into a small
// shared-memory ring slot (red_full mbarrier, count = kWarps). A dedicated REDUCER warp picks the slot up,
// adds the eight warp partials in fixed order and does everything that used to stall the whole CTA:
// partial-slot publish, device fence, per-row ticket, and — if it is the last contributor — the fixed-order
// cross-CTA combine and the row epilogue. The reducer has no data stores in flight, so its fence is cheap,
// and the consumers go straight on to their next stage: no bar.sync, no fence, no atomics on their path.
// (Measured before this change: making spans shorter — dynamic queue or oversubscribed grid — made K1oK2
// SLOWER, 34 vs 30 us, because each span end cost all 256 consumer threads a fence round trip.)
// ------------------------------------------------------------------------------------------------
constexpr int kRedSlots = 4;
constexpr int kRedStride = 4;   // floats per (slot, warp)

template <class Op>
__device__ __forceinline__ void reducer_segment(const typename Op::Params& p, const RowSched& s, const RowWorkspace& ws,
                                                long long row, bool row_began_before_span,
                                                const float (&part)[Op::K ? Op::K : 1]) {
    constexpr int K = Op::K;
    static_assert(K >= 1 && K <= 3, "slot holds 3 values");
    const int lane = threadIdx.x & 31;
    const typename Op::Row r = Op::row_begin(p, row);
    const long long rs = row * s.upr;
    const int first = span_owner(s, rs), last = span_owner(s, rs + s.upr - 1);
    double tot[K];
    if (first == last) {
#pragma unroll
        for (int k = 0; k < K; ++k) tot[k] = (double)part[k];
    } else {
        int is_last = 0;
        if (lane == 0) {
            st_slot(ws.partials + partial_slot(blockIdx.x, row_began_before_span) * kRowPartialStride, part[0],
                    K > 1 ? part[1] : 0.f, K > 2 ? part[2] : 0.f);
            __threadfence();
            unsigned int* counter = ws.counters + 1 + row;
            const unsigned int tk = atomicAdd(counter, 1u);
            is_last = (tk == (unsigned)(last - first));
            if (is_last) *counter = 0u;         // leave the workspace clean for the next launch
        }
        is_last = __shfl_sync(0xffffffffu, is_last, 0);
        if (!is_last) return;
        __threadfence();                        // acquire side
        double t[3] = {0.0, 0.0, 0.0};
        const int n = last - first + 1;
        for (int i = lane; i < n; i += 32) {    // contributors are the consecutive spans first..last
            const float4 v = ld_slot(ws.partials + partial_slot(first + i, i != 0) * kRowPartialStride);
            t[0] += (double)v.x; t[1] += (double)v.y; t[2] += (double)v.z;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) tot[k] = warp_sum(t[k]);
    }
    if (lane == 0) Op::row_end(p, r, row, tot);
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
//
// Warp roles: 0..7 consumers, 8 producer (one elected lane issues the bulk copies), 9 reducer.
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
    __host__ __device__ static constexpr int bytes() {
        return Op::kStages * stage_bytes() + 2 * Op::kStages * 8 + 2 * kRedSlots * 8 +
               kRedSlots * kWarps * kRedStride * 4 + 64;
    }
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
    uint64_t* red_full = empty + S;
    uint64_t* red_empty = red_full + kRedSlots;
    float* red_buf = reinterpret_cast<float*>(red_empty + kRedSlots);

    if (threadIdx.x == 0) {
        for (int i = 0; i < S; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, s.release_all ? kThreads : kWarps); }
        for (int i = 0; i < kRedSlots; ++i) { mbar_init(red_full + i, kWarps); mbar_init(red_empty + i, 1); }
        mbar_fence_init();
    }
    __syncthreads();

    long long u0, u1;
    cta_span(s, u0, u1);
    if (u0 >= u1) return;
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == kWarps) {
        // ---------------------------------------------------------------- producer (one elected thread)
        if (lane != 0) return;
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

    if (warp == kWarps + 1) {
        // ---------------------------------------------------------------- reducer warp
        if constexpr (Op::K > 0) {
            int rslot = 0;
            uint32_t rphase = 0;
            for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
                const RowSeg seg = row_segment(s, u0, u1, row);
                mbar_wait(red_full + rslot, rphase);
                // add the eight warp partials in fixed order (every lane computes the same values), then hand
                // the ring slot back. Both sides of this ring use the generic proxy (LDS here, STS in the
                // consumers), whose accesses the SM orders as issued, so releasing after the reads is enough.
                const float* sb = red_buf + rslot * kWarps * kRedStride;
                float part[Op::K];
#pragma unroll
                for (int k = 0; k < Op::K; ++k) {               // same order as block_sum: warp 0, 1, ... 7
                    float a = sb[k];
#pragma unroll
                    for (int w = 1; w < kWarps; ++w) a += sb[w * kRedStride + k];
                    part[k] = a;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(red_empty + rslot);
                if (++rslot == kRedSlots) { rslot = 0; rphase ^= 1u; }
                reducer_segment<Op>(p, s, ws, row, seg.begin > 0, part);
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers (warps 0..7)
    int stage = 0;
    uint32_t phase = 0;
    int rslot = 0;
    uint32_t rphase = 0;
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
                Op::unit(p, r, in, row * s.upr + u, acc);
            }
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
            // hand this warp's partial sums of the segment to the reducer warp and move on
            float v[Op::K];
#pragma unroll
            for (int k = 0; k < Op::K; ++k) v[k] = warp_sum(acc[k]);
            if (lane == 0) {
                mbar_wait(red_empty + rslot, rphase ^ 1u);
                float* dst = red_buf + (rslot * kWarps + warp) * kRedStride;
#pragma unroll
                for (int k = 0; k < Op::K; ++k) dst[k] = v[k];
                mbar_arrive(red_full + rslot);                   // release: publishes the stores above
            }
            __syncwarp();
            if (++rslot == kRedSlots) { rslot = 0; rphase ^= 1u; }
        }
    }
}

template <class Op>
static int launch_pipe(const typename Op::Params& p, RowWorkspace ws, long long B, long long D, int W, cudaStream_t st) {
    static bool configured[kMaxDevices] = {false};   // the opt-in is a per-device function attribute
    constexpr int smem = PipeSmem<Op>::bytes();
    const int slot = current_device_slot();
    if (!configured[slot] || slot == kMaxDevices - 1) {
        cudaError_t e = cudaFuncSetAttribute(pipe_row_kernel<Op>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        configured[slot] = true;
    }
    RowSched s = make_row_sched(B, D, W, Op::kOcc);
    // Oversubscription: `over` spans per resident CTA slot, span id == blockIdx.x, so the HARDWARE CTA scheduler
    // balances the load (a CTA that lands on a slow SM simply lets the others take more spans) and the
    // prologue / epilogue of one CTA overlaps with the streaming of the other CTAs resident on its SM.
    // Spans never shorter than 4 stages.
    static const int over = env_int("SISS_OVERSUB", 1);
    if (over > 1) {
        long long g = (long long)s.grid * over;
        const long long by_size = s.U / (4LL * kThreads);
        if (g > by_size) g = by_size;
        if (g > kMaxSpans) g = kMaxSpans;
        if (g > s.grid) { s.grid = (int)g; s.nspans = (int)g; }
    }
    static const int release_all = env_int("SISS_RELEASE_ALL", 1);   // see the release note in pipe_row_kernel
    s.release_all = release_all == 1 ? 1 : 0;
    pipe_row_kernel<Op><<<s.grid, kPipeThreads, smem, st>>>(p, ws, s);
    return (int)cudaGetLastError();
}

}  // namespace siss
