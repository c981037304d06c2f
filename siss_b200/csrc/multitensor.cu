// Multi-tensor K4: the same three-sum reduce (K4a) and scale-subtract-clip (K4b) as combine.cu, over a
// LIST of separately allocated gradient tensors — the layout the reference loop has (one tensor per
// UNet parameter, delete_celeb.py:717-750) — for callers that do not adopt GradCombiner's flat buffers.
// One launch for the whole list (the reference issues ~6 launches per parameter tensor, ~10^3 tensors).
//
// Device-side table, built once by the host (siss_b200.ops.MultiTensorPlan):
//   gx[i], ga[i], out[i]   pointers of tensor i            sizes[i]  elements of tensor i
//   chunk_prefix[i]        number of kMtChunk-element chunks in tensors 0..i-1   (chunk_prefix[n] = total)
// CTAs grid-stride over global chunk ids and binary-search the prefix table for the owning tensor.

#include "common.cuh"
#include "combine_scalars.cuh"

namespace siss {

int cached_sm_count();

constexpr int kMtChunk = 4096;      // elements per chunk: 256 threads x 4 float4
constexpr int kMtOcc = 4;
constexpr int kMtMaxGrid = 148 * 8;

struct MtTable {
    const float* const* gx;
    const float* const* ga;
    float* const* out;
    const long long* sizes;
    const long long* chunk_prefix;   // [n + 1]
    int n;
};

__device__ __forceinline__ int mt_find_tensor(const long long* __restrict__ prefix, int n, long long chunk) {
    int lo = 0, hi = n;              // invariant: prefix[lo] <= chunk < prefix[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= chunk) lo = mid; else hi = mid;
    }
    return lo;
}

struct MtNormWorkspace { unsigned int* counter; double* partials; };

__global__ void __launch_bounds__(kThreads, kMtOcc)
mt_norm3_kernel(MtTable tb, double* __restrict__ sums3, MtNormWorkspace ws) {
    __shared__ double red[3 * kWarps];
    __shared__ int flag;
    double acc[3] = {0.0, 0.0, 0.0};
    const long long total_chunks = tb.chunk_prefix[tb.n];
    for (long long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        const int t = mt_find_tensor(tb.chunk_prefix, tb.n, c);
        const long long off = (c - tb.chunk_prefix[t]) * kMtChunk;
        const long long len = min((long long)kMtChunk, tb.sizes[t] - off);
        const float* __restrict__ x = tb.gx[t] + off;
        const float* __restrict__ a = tb.ga[t] + off;
        if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a)) & 15u) == 0) {
            const int nv = (int)(len / 4);
            uint4 rx[4], ra[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = j * kThreads + threadIdx.x;
                ok[j] = i < nv;
                if (ok[j]) { rx[j] = ldg_stream(x + 4 * i); ra[j] = ldg_stream(a + 4 * i); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!ok[j]) continue;
                float xf[4], af[4];
                VecTraits<float>::unpack(rx[j], xf);
                VecTraits<float>::unpack(ra[j], af);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const double xd = (double)xf[q], ad = (double)af[q];
                    acc[0] = fma(xd, xd, acc[0]); acc[1] = fma(ad, ad, acc[1]); acc[2] = fma(xd, ad, acc[2]);
                }
            }
            for (long long i = (long long)nv * 4 + threadIdx.x; i < len; i += kThreads) {
                const double xd = (double)x[i], ad = (double)a[i];
                acc[0] = fma(xd, xd, acc[0]); acc[1] = fma(ad, ad, acc[1]); acc[2] = fma(xd, ad, acc[2]);
            }
        } else {
            for (long long i = threadIdx.x; i < len; i += kThreads) {
                const double xd = (double)x[i], ad = (double)a[i];
                acc[0] = fma(xd, xd, acc[0]); acc[1] = fma(ad, ad, acc[1]); acc[2] = fma(xd, ad, acc[2]);
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
            if (threadIdx.x == 0) { sums3[0] = t[0]; sums3[1] = t[1]; sums3[2] = t[2]; }
        }
    }
}

__global__ void __launch_bounds__(kThreads, kMtOcc)
mt_combine_kernel(MtTable tb, const double* __restrict__ sums3, int mode, float value, float max_norm, int inf_guard,
                  float* __restrict__ stats5) {
    const CombineScalars cs = combine_scalars_from(sums3[0], sums3[1], sums3[2], mode, value, max_norm, inf_guard, stats5,
                                                   blockIdx.x == 0 && threadIdx.x == 0);
    const float s = cs.s, clip = cs.clip;
    const long long total_chunks = tb.chunk_prefix[tb.n];
    for (long long c = blockIdx.x; c < total_chunks; c += gridDim.x) {
        const int t = mt_find_tensor(tb.chunk_prefix, tb.n, c);
        const long long off = (c - tb.chunk_prefix[t]) * kMtChunk;
        const long long len = min((long long)kMtChunk, tb.sizes[t] - off);
        const float* x = tb.gx[t] + off;
        const float* a = tb.ga[t] + off;
        float* o = tb.out[t] + off;
        if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(o)) & 15u) == 0) {
            const int nv = (int)(len / 4);
            uint4 rx[4], ra[4];
            bool ok[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = j * kThreads + threadIdx.x;
                ok[j] = i < nv;
                if (ok[j]) { rx[j] = ldg_v4(x + 4 * i); ra[j] = ldg_v4(a + 4 * i); }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!ok[j]) continue;
                const int i = j * kThreads + threadIdx.x;
                float xf[4], af[4], of[4];
                VecTraits<float>::unpack(rx[j], xf);
                VecTraits<float>::unpack(ra[j], af);
#pragma unroll
                for (int q = 0; q < 4; ++q) of[q] = __fmul_rn(__fsub_rn(xf[q], __fmul_rn(s, af[q])), clip);
                stg_stream(o + 4 * i, VecTraits<float>::pack(of));
            }
            for (long long i = (long long)nv * 4 + threadIdx.x; i < len; i += kThreads)
                o[i] = __fmul_rn(__fsub_rn(x[i], __fmul_rn(s, a[i])), clip);
        } else {
            for (long long i = threadIdx.x; i < len; i += kThreads)
                o[i] = __fmul_rn(__fsub_rn(x[i], __fmul_rn(s, a[i])), clip);
        }
    }
}

static int mt_grid(long long total_chunks) {
    long long grid = (long long)cached_sm_count() * kMtOcc;
    if (grid > kMtMaxGrid) grid = kMtMaxGrid;
    if (total_chunks < grid) grid = total_chunks;
    if (grid < 1) grid = 1;
    return (int)grid;
}

}  // namespace siss

using namespace siss;

extern "C" {

int siss_mt_chunk_elems(void) { return kMtChunk; }

int siss_mt_norm3(const float* const* d_gx, const float* const* d_ga, const int64_t* d_sizes,
                  const int64_t* d_chunk_prefix, int n_tensors, int64_t total_chunks, double* sums3,
                  void* workspace, siss_stream_t stream) {
    if (!d_gx || !d_ga || !d_sizes || !d_chunk_prefix || !sums3 || !workspace || n_tensors < 1 || total_chunks < 0)
        return SISS_EINVAL;
    MtTable tb{d_gx, d_ga, nullptr, (const long long*)d_sizes, (const long long*)d_chunk_prefix, n_tensors};
    MtNormWorkspace ws{reinterpret_cast<unsigned int*>(workspace),
                       reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + 256)};
    mt_norm3_kernel<<<mt_grid(total_chunks), kThreads, 0, (cudaStream_t)stream>>>(tb, sums3, ws);
    return (int)cudaGetLastError();
}

int siss_mt_combine(const float* const* d_gx, const float* const* d_ga, float* const* d_out, const int64_t* d_sizes,
                    const int64_t* d_chunk_prefix, int n_tensors, int64_t total_chunks, const double* sums3,
                    int mode, float value, float max_norm, int inf_guard, float* stats5, siss_stream_t stream) {
    if (!d_gx || !d_ga || !d_out || !d_sizes || !d_chunk_prefix || !sums3 || n_tensors < 1 || total_chunks < 0)
        return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_NONE) return SISS_EINVAL;
    MtTable tb{d_gx, d_ga, d_out, (const long long*)d_sizes, (const long long*)d_chunk_prefix, n_tensors};
    mt_combine_kernel<<<mt_grid(total_chunks), kThreads, 0, (cudaStream_t)stream>>>(tb, sums3, mode, value, max_norm,
                                                                                    inf_guard, stats5);
    return (int)cudaGetLastError();
}

}  // extern "C"
