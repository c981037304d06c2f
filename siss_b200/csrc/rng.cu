// Opt-in device-side draws: the noise tensor and the per-row timestep / Bernoulli mask
// from the counter-based stream of philox.cuh, replacing torch.randn (delete_celeb.py:581), torch.randint (:593)
// and the CPU torch.rand(B) > lambd + its H2D copy (losses/ddpm_deletion_loss.py:18). Values depend only on
// (seed, draw, global index), so N data-parallel ranks that pass their global offsets draw exactly the slices of
// the 1-rank tensors — the reference's "same seed on every rank" defect cannot occur.

#include "philox.cuh"

namespace siss {

int cached_sm_count();

template <typename T>
__global__ void __launch_bounds__(kThreads, 4)
randn_kernel(T* __restrict__ out, long long n, long long nvec, RngStream s, const unsigned long long* __restrict__ d_draw,
             unsigned long long elem_offset) {
    constexpr int W = VecTraits<T>::N;
    rng_draw_from_device(s, d_draw);
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long u = (long long)blockIdx.x * kThreads + threadIdx.x; u < nvec; u += stride) {
        float z[W];
        rng_normals<W>(s, elem_offset + (unsigned long long)u * W, z);
        stg_stream(out + u * W, VecTraits<T>::pack(z));
    }
    for (long long i = nvec * W + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        float z[1];
        rng_normals<1>(s, elem_offset + (unsigned long long)i, z);
        VecTraits<T>::store1(out + i, z[0]);
    }
}

__global__ void draw_rows_kernel(long long B, RngStream s, const unsigned long long* __restrict__ d_draw,
                                 unsigned long long row_offset, long long t_lo,
                                 unsigned int t_span, float lambd, int64_t* __restrict__ ts, uint8_t* __restrict__ keep) {
    rng_draw_from_device(s, d_draw);
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= B) return;
    uint32_t w[4];
    philox4x32_10(s, row_offset + (unsigned long long)r, w);
    if (ts) ts[r] = t_lo + (long long)(w[0] % t_span);
    if (keep) keep[r] = rng_uniform(w[1]) > lambd ? 1 : 0;     // torch.rand(B) > lambd
}

template <typename T>
static int launch_randn(void* out, long long n, uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
                        cudaStream_t st) {
    constexpr int W = VecTraits<T>::N;
    const bool vec = aligned16(out) && (elem_offset % 4 == 0);
    const long long nvec = vec ? n / W : 0;
    long long work = (nvec > n - nvec * W ? nvec : n - nvec * W);
    long long grid = (work + kThreads - 1) / kThreads;
    const long long cap = (long long)cached_sm_count() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    randn_kernel<T><<<(int)grid, kThreads, 0, st>>>((T*)out, n, nvec, make_rng_stream(seed, draw, kRngNoise),
                                                    (const unsigned long long*)d_draw, elem_offset);
    return (int)cudaGetLastError();
}

}  // namespace siss

using namespace siss;

extern "C" {

int siss_randn(void* out, int64_t n, int dtype, uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
               siss_stream_t stream) {
    if (!out || n < 0 || (draw >> 62)) return SISS_EINVAL;
    if (n == 0) return SISS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case SISS_F32:  return launch_randn<float>(out, n, seed, draw, d_draw, elem_offset, st);
        case SISS_BF16: return launch_randn<__nv_bfloat16>(out, n, seed, draw, d_draw, elem_offset, st);
        case SISS_F16:  return launch_randn<__half>(out, n, seed, draw, d_draw, elem_offset, st);
        default: return SISS_EUNSUPPORTED;
    }
}

int siss_draw_rows(int64_t* timesteps, uint8_t* keep_mask, int64_t B, uint64_t seed, uint64_t draw, const uint64_t* d_draw,
                   uint64_t row_offset, int64_t t_lo, int64_t t_hi, double lambd, siss_stream_t stream) {
    if ((!timesteps && !keep_mask) || B < 0 || (draw >> 62)) return SISS_EINVAL;
    if (timesteps && (t_lo < 0 || t_hi <= t_lo || t_hi - t_lo > 0x7FFFFFFFLL)) return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    const unsigned int span = timesteps ? (unsigned int)(t_hi - t_lo) : 1u;
    draw_rows_kernel<<<(int)((B + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        B, make_rng_stream(seed, draw, kRngRows), (const unsigned long long*)d_draw, row_offset, t_lo, span, (float)lambd,
        timesteps, keep_mask);
    return (int)cudaGetLastError();
}

}  // extern "C"
