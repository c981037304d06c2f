// Scalar logic of K4b (siss_combine): scaling factor per mode, norm of the combination from the three
// sums, clip coefficient — in the reference's fp32 op order. Shared by combine.cu and p2p.cu.
#pragma once

#include "common.cuh"

namespace siss {

struct CombineScalars {
    float s;     // scaling factor
    float clip;  // clip coefficient (<= 1)
};

// Evaluated redundantly by every thread (a handful of scalar ops); fp32 op order of the reference.
__device__ __forceinline__ CombineScalars combine_scalars_from(double sxx, double saa, double sxa, int mode, float value,
                                                               float max_norm, int inf_guard, float* stats5,
                                                               bool write_stats) {
    const float n_x = sqrtf((float)sxx);  // torch.sqrt(sum of per-tensor norm**2), delete_celeb.py:733-734
    const float n_a = sqrtf((float)saa);
    float s;
    if (mode == SISS_COMBINE_NONE) {
        s = 0.0f;
    } else if (mode == SISS_COMBINE_ERASEDIFF) {
        // eta - <g_x,g_a> / ||g_a||**2 ; -max(., 0)      delete_celeb.py:741-742
        float sf = __fsub_rn(value, __fdiv_rn((float)sxa, __fmul_rn(n_a, n_a)));
        sf = (0.0f > sf) ? 0.0f : sf;  // python max(sf, 0): keeps NaN
        s = -sf;
    } else {
        s = __fdiv_rn(value, n_a);     // scaling_norm / ||g_a||   delete_celeb.py:746
        if (inf_guard && isinf(s)) s = 0.0f;  // delete_tshirt.py:688-690
    }
    // ||g_x - s g_a||^2 from the three sums, in fp64
    const double sd = (double)s;
    double tn2 = sxx - 2.0 * sd * sxa + sd * sd * saa;
    if (tn2 < 0.0) tn2 = 0.0;
    const float tn = (float)sqrt(tn2);
    float clip = 1.0f;
    if (max_norm > 0.0f) {
        // torch.nn.utils.clip_grad_norm_: max_norm / (total_norm + 1e-6), clamped to 1
        clip = __fdiv_rn(max_norm, __fadd_rn(tn, 1e-6f));
        clip = (clip > 1.0f) ? 1.0f : clip;
    }
    if (write_stats && stats5) {
        stats5[0] = n_x; stats5[1] = n_a; stats5[2] = s; stats5[3] = tn; stats5[4] = clip;
    }
    CombineScalars r; r.s = s; r.clip = clip;
    return r;
}

}  // namespace siss
