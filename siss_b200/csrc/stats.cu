// Fused statistics epilogue: the 16 per-batch logging scalars of delete_celeb.py:626-656 in ONE
// launch, computed from the O(B) per-sample sums K2/K3 already produced.
// The reference does up to 20 `.item()` host syncs over [B,C,H,W] tensors here; this kernel reads
// 4*B floats and writes 16, and the host reads them back whenever it logs.
//
//   out[0..3]   loss_x / mean, max, min, std      (mean over all B*D elements; max/min/std over the
//   out[4..7]   loss_a / mean, max, min, std       per-sample means row/D; std unbiased like torch.std)
//   out[8..11]  importance_weight_x / mean, max, min, std
//   out[12..15] importance_weight_a / mean, max, min, std
// A NULL input leaves its four outputs NaN. B == 1 gives std = NaN, as torch does.

#include <math.h>

#include "common.cuh"

namespace siss {

// fixed-order block reduction of (sum, max, min) in double/float
__device__ void stat4(const float* __restrict__ v, long long B, double inv_scale, float* __restrict__ out4,
                      double* smem) {
    const int tid = threadIdx.x;
    if (v == nullptr) {
        if (tid < 4) out4[tid] = nanf("");
        return;
    }
    double s = 0.0;
    float mx = -INFINITY, mn = INFINITY;
    bool nan_seen = false;
    for (long long i = tid; i < B; i += kThreads) {
        const float x = (float)((double)v[i] * inv_scale);
        // per-sample mean as eager forms it: fp32 sum / D — the row sum is already fp32
        s += (double)x;
        nan_seen |= (x != x);
        mx = fmaxf(mx, x);
        mn = fminf(mn, x);
    }
    double* ss = smem;                                          // [kThreads]
    float* smx = reinterpret_cast<float*>(smem + kThreads);     // [kThreads]
    float* smn = smx + kThreads;                                // [kThreads]
    int* snan = reinterpret_cast<int*>(smn + kThreads);
    ss[tid] = s; smx[tid] = mx; smn[tid] = mn; snan[tid] = nan_seen;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (tid < o) {
            ss[tid] += ss[tid + o];
            smx[tid] = fmaxf(smx[tid], smx[tid + o]);
            smn[tid] = fminf(smn[tid], smn[tid + o]);
            snan[tid] |= snan[tid + o];
        }
        __syncthreads();
    }
    const double mean = ss[0] / (double)B;
    const float fmx = smx[0], fmn = smn[0];
    const bool any_nan = snan[0] != 0;
    __syncthreads();
    double q = 0.0;
    for (long long i = tid; i < B; i += kThreads) {
        const double d = (double)(float)((double)v[i] * inv_scale) - mean;
        q += d * d;
    }
    ss[tid] = q;
    __syncthreads();
    for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (tid < o) ss[tid] += ss[tid + o];
        __syncthreads();
    }
    if (tid == 0) {
        const float nanv = nanf("");
        out4[0] = (float)mean;
        out4[1] = any_nan ? nanv : fmx;   // torch.max / min propagate NaN
        out4[2] = any_nan ? nanv : fmn;
        out4[3] = (B > 1) ? (float)sqrt(ss[0] / (double)(B - 1)) : nanv;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kThreads)
batch_stats_kernel(const float* row_loss_x, const float* row_loss_a, const float* w_x, const float* w_a,
                   long long B, long long D, float* __restrict__ out16) {
    __shared__ double smem[kThreads + kThreads + kThreads / 2 + kThreads / 2 + 8];
    const double inv_d = 1.0 / (double)D;
    stat4(row_loss_x, B, inv_d, out16 + 0, smem);
    stat4(row_loss_a, B, inv_d, out16 + 4, smem);
    stat4(w_x, B, 1.0, out16 + 8, smem);
    stat4(w_a, B, 1.0, out16 + 12, smem);
}

}  // namespace siss

extern "C" int siss_batch_stats(const float* row_loss_x, const float* row_loss_a, const float* w_x, const float* w_a,
                                int64_t B, int64_t D, float* out16, siss_stream_t stream) {
    if (!out16 || B < 1 || D < 1) return SISS_EINVAL;
    siss::batch_stats_kernel<<<1, siss::kThreads, 0, (cudaStream_t)stream>>>(row_loss_x, row_loss_a, w_x, w_a, B, D, out16);
    return (int)cudaGetLastError();
}
