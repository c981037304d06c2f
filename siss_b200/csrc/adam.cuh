// AdamW / EMA element updates and their scalars, shared by the flat fused optimiser (optim.cu) and the
// peer-memory sharded optimiser (p2p.cu). Update rule in torch.optim.AdamW's single-tensor op order (fp32):
//     p *= 1 - lr*wd ; m += (1-b1) (g - m) ; v = v*b2 + (1-b2) g g
//     p += (-lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
#pragma once

#include <math.h>

#include "common.cuh"

namespace siss {

struct AdamScalars {
    float decay;        // 1 - lr * weight_decay
    float w1;           // 1 - beta1
    float beta2;
    float w2;           // 1 - beta2
    float bc2_sqrt;     // sqrt(1 - beta2^t)
    float neg_step;     // -lr / (1 - beta1^t)
    float eps;
    double lr, beta1_d, beta2_d;   // for the device-side step counter variant
    double wd;                     // for the device-side lr variant
    float ema_omd;                 // 1 - ema_decay
};

// Bias corrections from a step count held in DEVICE memory (so a captured CUDA graph stays valid from one
// optimiser step to the next): same double-precision expressions the host path evaluates.
__device__ __forceinline__ void adam_bias_from_step(AdamScalars& a, long long step) {
    const double bc1 = 1.0 - pow(a.beta1_d, (double)step), bc2 = 1.0 - pow(a.beta2_d, (double)step);
    a.bc2_sqrt = (float)sqrt(bc2);
    a.neg_step = (float)(-(a.lr / bc1));
}

// lr / ema_decay from device memory; must run BEFORE adam_bias_from_step (neg_step uses a.lr)
__device__ __forceinline__ void adam_sched_from_device(AdamScalars& a, const double* d_sched, long long host_step) {
    a.lr = d_sched[0];
    a.decay = (float)(1.0 - a.lr * a.wd);
    a.ema_omd = (float)(1.0 - d_sched[1]);
    adam_bias_from_step(a, host_step);
}

__device__ __forceinline__ float ema_update(float shadow, float p, float omd) {
    return __fsub_rn(shadow, __fmul_rn(__fsub_rn(shadow, p), omd));   // s.sub_(omd * (s - p))
}

__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, const AdamScalars& a) {
    p = __fmul_rn(p, a.decay);
    m = __fadd_rn(m, __fmul_rn(a.w1, __fsub_rn(g, m)));                        // lerp, small weight branch
    v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.w2, g), g));    // mul_ ; addcmul_
    const float denom = __fadd_rn(__fdiv_rn(sqrtf(v), a.bc2_sqrt), a.eps);
    p = __fadd_rn(p, __fmul_rn(a.neg_step, __fdiv_rn(m, denom)));              // addcdiv_
}

// python-float scalars of torch.optim.AdamW's single-tensor path, rounded to fp32 where the tensor op does
inline AdamScalars make_adam_scalars(double lr, double beta1, double beta2, double eps, double weight_decay,
                                     long long step, double ema_decay, long long& host_step) {
    AdamScalars as;
    as.decay = (float)(1.0 - lr * weight_decay);
    as.w1 = (float)(1.0 - beta1);
    as.beta2 = (float)beta2;
    as.w2 = (float)(1.0 - beta2);
    const double hstep = (double)(step < 1 ? 1 : step);
    const double bc1 = 1.0 - pow(beta1, hstep), bc2 = 1.0 - pow(beta2, hstep);
    as.bc2_sqrt = (float)sqrt(bc2);
    as.neg_step = (float)(-(lr / bc1));
    as.eps = (float)eps;
    as.lr = lr; as.beta1_d = beta1; as.beta2_d = beta2;
    as.wd = weight_decay;
    as.ema_omd = (float)(1.0 - ema_decay);
    host_step = (long long)hstep;
    return as;
}

}  // namespace siss
