// Forward-diffusion element arithmetic shared by K1, K1oK2 and the membership-metric kernels.
#pragma once

#include "common.cuh"

namespace siss {

// sqrt(abar_t) and sqrt(1 - abar_t) the way diffusers 0.27.2 DDPMScheduler.add_noise forms them:
// the table is cast to the sample dtype first, and each op result is rounded to that dtype.
template <typename T>
__device__ __forceinline__ void noise_coeffs(const float* __restrict__ ac, int t, float& sa, float& s1) {
    using VT = VecTraits<T>;
    const float a = VT::round(ac[t]);
    sa = VT::round(sqrtf(a));
    s1 = VT::round(sqrtf(VT::round(__fsub_rn(1.0f, a))));
}

// x_t element: round(round(sa*x) + round(s1*n)); __f*_rn blocks FMA contraction so the fp32 path
// is bit-identical to eager's mul, mul, add.
template <typename T>
__device__ __forceinline__ float noised(float sa, float s1, float x, float n) {
    using VT = VecTraits<T>;
    return VT::round(__fadd_rn(VT::round(__fmul_rn(sa, x)), VT::round(__fmul_rn(s1, n))));
}

}  // namespace siss
