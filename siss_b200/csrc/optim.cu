// K4b fused with the optimiser step: the combined, clipped gradient
//     g = clip * (G_x - s * G_a)
// is consumed in registers by a decoupled-weight-decay Adam update of the flat fp32 parameter buffer
// (torch.optim.AdamW as configured at config/delete_celeb.yaml:127-134, stepped at delete_celeb.py:769)
// and G_x / G_a are cleared in the same pass, so `optimizer.zero_grad()` (:773) and the two memsets of
// the unfused path disappear too:
//     unfused: K4b 12 + AdamW 28 + 2 memsets 8 = 48 B/param, 4 launches
//     fused  : read G_x G_a p m v (20) + write p m v G_x G_a (20) = 40 B/param, 1 launch
// Update rule, in torch's single-tensor op order (fp32):
//     p *= 1 - lr*wd ; m += (1-b1) (g - m) ; v = v*b2 + (1-b2) g g
//     p += (-lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
// Optional, same pass (+8 B/param): the EMA shadow update the reference runs right after the optimiser step
// (`ema_model.step(unet.parameters())`, delete_celeb.py:776-777; diffusers EMAModel.step):
//     shadow -= (1 - decay) * (shadow - p_new)
// Optional `d_sched` = {lr, ema_decay} in DEVICE memory: learning-rate schedules (lr_scheduler.step(), :770) and
// EMA warm-up then work under CUDA-graph replay, like the device-side step counter.

#include "common.cuh"
#include "combine_scalars.cuh"
#include "adam.cuh"

namespace siss {

int cached_sm_count();

constexpr int kOptOcc = 3;
constexpr int kOptUnroll = 2;

__global__ void counter_add_kernel(long long* p, long long v) { *p += v; }

template <bool TWO_TERM, bool EMA>
__global__ void __launch_bounds__(kThreads, kOptOcc)
combine_adamw_kernel(float* __restrict__ gx, float* __restrict__ ga, long long n, long long nvec,
                     const double* __restrict__ sums3, int mode, float value, float max_norm, int inf_guard,
                     float* __restrict__ param, float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                     AdamScalars as, long long host_step, const long long* __restrict__ d_step,
                     const double* __restrict__ d_sched, float* __restrict__ ema, int zero_grads,
                     float* __restrict__ grad_out, float* __restrict__ stats5) {
    if (d_sched != nullptr) adam_sched_from_device(as, d_sched, host_step);   // lr / ema decay from device memory
    if (d_step != nullptr) adam_bias_from_step(as, *d_step);   // step count from device memory (graph replay)
    float s = 0.f, clip = 1.f;
    if (sums3 != nullptr) {
        const CombineScalars cs = combine_scalars_from(sums3[0], sums3[1], sums3[2], mode, value, max_norm, inf_guard,
                                                       stats5, blockIdx.x == 0 && threadIdx.x == 0);
        s = cs.s; clip = cs.clip;
    }
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
    const long long chunk = (long long)kThreads * kOptUnroll;
    const long long nchunks = (nvec + chunk - 1) / chunk;
    for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
        const long long base = (nchunks - 1 - c) * chunk + threadIdx.x;   // reverse of K4a's walk (L2 tail reuse)
        uint4 rx[kOptUnroll], ra[kOptUnroll], rp[kOptUnroll], rm[kOptUnroll], rv[kOptUnroll], re[kOptUnroll];
        bool ok[kOptUnroll];
#pragma unroll
        for (int j = 0; j < kOptUnroll; ++j) {
            const long long i = base + (long long)j * kThreads;
            ok[j] = i < nvec;
            if (ok[j]) {
                rx[j] = ldg_v4(gx + 4 * i);
                if (TWO_TERM) ra[j] = ldg_v4(ga + 4 * i);
                rp[j] = ldg_v4(param + 4 * i);
                rm[j] = ldg_v4(exp_avg + 4 * i);
                rv[j] = ldg_v4(exp_avg_sq + 4 * i);
                if (EMA) re[j] = ldg_v4(ema + 4 * i);
            }
        }
#pragma unroll
        for (int j = 0; j < kOptUnroll; ++j) {
            if (!ok[j]) continue;
            const long long i = base + (long long)j * kThreads;
            float x[4], a[4], p[4], m[4], v[4], g[4];
            VecTraits<float>::unpack(rx[j], x);
            if (TWO_TERM) VecTraits<float>::unpack(ra[j], a);
            VecTraits<float>::unpack(rp[j], p);
            VecTraits<float>::unpack(rm[j], m);
            VecTraits<float>::unpack(rv[j], v);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                g[q] = TWO_TERM ? __fmul_rn(__fsub_rn(x[q], __fmul_rn(s, a[q])), clip) : __fmul_rn(x[q], clip);
                adam_update(g[q], p[q], m[q], v[q], as);
            }
            stg_stream(param + 4 * i, VecTraits<float>::pack(p));
            stg_stream(exp_avg + 4 * i, VecTraits<float>::pack(m));
            stg_stream(exp_avg_sq + 4 * i, VecTraits<float>::pack(v));
            if (EMA) {
                float e[4];
                VecTraits<float>::unpack(re[j], e);
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = ema_update(e[q], p[q], as.ema_omd);
                stg_stream(ema + 4 * i, VecTraits<float>::pack(e));
            }
            if (grad_out) stg_stream(grad_out + 4 * i, VecTraits<float>::pack(g));
            if (zero_grads) {
                if (grad_out != gx) stg_stream(gx + 4 * i, zero4);
                if (TWO_TERM) stg_stream(ga + 4 * i, zero4);
            }
        }
    }
    {   // scalar remainder / unaligned buffers
        const long long start = nvec * 4;
        const long long stride = (long long)gridDim.x * kThreads;
        for (long long i = start + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
            const float g = TWO_TERM ? __fmul_rn(__fsub_rn(gx[i], __fmul_rn(s, ga[i])), clip) : __fmul_rn(gx[i], clip);
            float p = param[i], m = exp_avg[i], v = exp_avg_sq[i];
            adam_update(g, p, m, v, as);
            param[i] = p; exp_avg[i] = m; exp_avg_sq[i] = v;
            if (EMA) ema[i] = ema_update(ema[i], p, as.ema_omd);
            if (grad_out) grad_out[i] = g;
            if (zero_grads) {
                if (grad_out != gx) gx[i] = 0.f;
                if (TWO_TERM) ga[i] = 0.f;
            }
        }
    }
}

}  // namespace siss

using namespace siss;

extern "C" int siss_counter_add(int64_t* d_counter, int64_t value, siss_stream_t stream) {
    if (!d_counter) return SISS_EINVAL;
    counter_add_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((long long*)d_counter, (long long)value);
    return (int)cudaGetLastError();
}

extern "C" int siss_combine_adamw(float* g_x, float* g_a, int64_t n, const double* sums3, int mode, float value,
                                  float max_norm, int inf_guard, float* param, float* exp_avg, float* exp_avg_sq,
                                  double lr, double beta1, double beta2, double eps, double weight_decay, int64_t step,
                                  const int64_t* d_step, const double* d_sched, float* ema_param, double ema_decay,
                                  int zero_grads, float* grad_out, float* stats5, siss_stream_t stream) {
    if (!g_x || !param || !exp_avg || !exp_avg_sq || n < 0 || (step < 1 && !d_step)) return SISS_EINVAL;
    if (ema_param && !d_sched && !(ema_decay >= 0.0 && ema_decay <= 1.0)) return SISS_EINVAL;
    if (mode < SISS_COMBINE_SCALING_NORM || mode > SISS_COMBINE_NONE) return SISS_EINVAL;
    const bool two_term = (g_a != nullptr) && mode != SISS_COMBINE_NONE;
    if (mode != SISS_COMBINE_NONE && (!g_a || !sums3)) return SISS_EINVAL;
    long long hs;
    const AdamScalars as = make_adam_scalars(lr, beta1, beta2, eps, weight_decay, step, ema_decay, hs);
    const long long* dstep = (const long long*)d_step;
    bool al = aligned16(g_x) && aligned16(param) && aligned16(exp_avg) && aligned16(exp_avg_sq) && aligned16(grad_out) &&
              aligned16(ema_param);
    if (two_term) al = al && aligned16(g_a);
    const long long nvec = al ? n / 4 : 0;
    const long long chunk = (long long)kThreads * kOptUnroll;
    long long work = (nvec + chunk - 1) / chunk;
    if (nvec * 4 < n) {
        const long long tail_blocks = (n - nvec * 4 + kThreads - 1) / kThreads;
        if (tail_blocks > work) work = tail_blocks;
    }
    long long grid = (long long)cached_sm_count() * kOptOcc;
    if (work < grid) grid = work;
    if (grid < 1) grid = 1;
    cudaStream_t st = (cudaStream_t)stream;
#define SISS_LAUNCH_ADAMW(TT, EM)                                                                                         \
    combine_adamw_kernel<TT, EM><<<(int)grid, kThreads, 0, st>>>(g_x, g_a, n, nvec, sums3, mode, value, max_norm, inf_guard,  \
                                                                 param, exp_avg, exp_avg_sq, as, hs, dstep, d_sched,          \
                                                                 ema_param, zero_grads, grad_out, stats5)
    if (two_term) { if (ema_param) SISS_LAUNCH_ADAMW(true, true); else SISS_LAUNCH_ADAMW(true, false); }
    else          { if (ema_param) SISS_LAUNCH_ADAMW(false, true); else SISS_LAUNCH_ADAMW(false, false); }
#undef SISS_LAUNCH_ADAMW
    return (int)cudaGetLastError();
}
