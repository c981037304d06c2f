// Membership-loss metric kernels (reference metrics/class_membership.py:66-116).
//
// The metric evaluates, for I sampled images x N_n shared noise draws at one timestep t,
//     all_loss[r]      = sum_{chw} (unet(add_noise(x0[i], eps[j], t))[r] - eps[j])^2,    r = i * N_n + j
// (and the same for the deletion images), then the mean over r. The reference materialises the expanded
// [I*N_n, C, H, W] image and noise tensors (:76-86), runs add_noise twice over them (:92-93) and, per
// eval batch, forms (pred - noise)^2 and reduces it (:108-109). Here the expansion is an index map:
//   * siss_membership_add_noise : both noisy batches for expanded rows [row0, row0 + rows) straight from the
//     [I, D] images and the [N_n, D] noise (image row r / N_n, noise row r % N_n) — 3 reads + 2 writes per
//     element, nothing expanded, same rounding sequence as K1 (noise.cuh);
//   * siss_membership_sqerr     : both per-row sums of squared errors in one pass over pred_x, pred_a and the
//     noise rows (3 reads per element, O(rows) writes), rows split over CTAs like K2 / K3 (rowtile.cuh).
// Eval-time code: LDG path only (no TMA ring), 128-bit accesses, persistent grid.

#include "rowtile.cuh"
#include "noise.cuh"

namespace siss {

constexpr int kMbVpt = 2;
constexpr int kMbOcc = 4;

template <typename T, int W>
__global__ void __launch_bounds__(kThreads, kMbOcc)
membership_add_noise_kernel(const T* __restrict__ x0, const T* __restrict__ a0, const T* __restrict__ noise,
                            const float* __restrict__ ac, int t, T* __restrict__ xt_x, T* __restrict__ xt_a,
                            long long row0, long long n_noise, RowSched s) {
    constexpr int VPT = kMbVpt;
    long long u0, u1;
    cta_span(s, u0, u1);
    if (u0 >= u1) return;
    float sa, s1;
    noise_coeffs<T>(ac, t, sa, s1);     // one timestep for the whole call (:89)
    for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
        const RowSeg sg = row_segment(s, u0, u1, row);
        const long long r = row0 + row;
        const long long img_off = (r / n_noise) * s.D, noise_off = (r % n_noise) * s.D, out_off = row * s.D;
        for (long long ub = sg.begin; ub < sg.end; ub += (long long)kThreads * VPT) {
            RawUnit<T, W> rx[VPT], ra[VPT], rn[VPT];
            long long e[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < sg.end) ? u * W : -1;
                if (e[j] >= 0) {
                    fetch_raw<T, W>(x0 + img_off + e[j], rx[j]);
                    fetch_raw<T, W>(a0 + img_off + e[j], ra[j]);
                    fetch_raw<T, W>(noise + noise_off + e[j], rn[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (e[j] < 0) continue;
                float x[W], n[W], o[W];
                decode_raw<T, W>(rn[j], n);
                decode_raw<T, W>(rx[j], x);
#pragma unroll
                for (int k = 0; k < W; ++k) o[k] = noised<T>(sa, s1, x[k], n[k]);
                store_unit<T, W>(xt_x + out_off + e[j], o);
                decode_raw<T, W>(ra[j], x);
#pragma unroll
                for (int k = 0; k < W; ++k) o[k] = noised<T>(sa, s1, x[k], n[k]);
                store_unit<T, W>(xt_a + out_off + e[j], o);
            }
        }
    }
}

// W elements of an fp32 prediction next to one unit (W elements) of the T-typed noise
template <int W>
__device__ __forceinline__ void fetch_pred(const float* p, uint4 (&r)[(W + 3) / 4], float& s) {
    if constexpr (W == 1) s = __ldg(p);
    else {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) r[i] = ldg_stream(p + 4 * i);
    }
}
template <int W>
__device__ __forceinline__ void decode_pred(const uint4 (&r)[(W + 3) / 4], float s, float (&f)[W]) {
    if constexpr (W == 1) f[0] = s;
    else {
#pragma unroll
        for (int i = 0; i < W / 4; ++i) {
            float tmp[4];
            VecTraits<float>::unpack(r[i], tmp);
#pragma unroll
            for (int q = 0; q < 4; ++q) f[4 * i + q] = tmp[q];
        }
    }
}

template <typename T, int W>
__global__ void __launch_bounds__(kThreads, kMbOcc)
membership_sqerr_kernel(const float* __restrict__ pred_x, const float* __restrict__ pred_a,
                        const T* __restrict__ noise, float* __restrict__ sum_x, float* __restrict__ sum_a,
                        long long row0, long long n_noise, RowWorkspace ws, RowSched s) {
    __shared__ float red[2 * kWarps];
    __shared__ int flag;
    long long u0, u1;
    cta_span(s, u0, u1);
    if (u0 >= u1) return;
    for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
        const RowSeg sg = row_segment(s, u0, u1, row);
        const long long noise_off = ((row0 + row) % n_noise) * s.D, out_off = row * s.D;
        float acc[2] = {0.f, 0.f};
        for (long long ub = sg.begin; ub < sg.end; ub += (long long)kThreads * kMbVpt) {
            uint4 px[kMbVpt][(W + 3) / 4], pa[kMbVpt][(W + 3) / 4];
            float sx[kMbVpt], sa[kMbVpt];
            RawUnit<T, W> rn[kMbVpt];
            long long e[kMbVpt];
#pragma unroll
            for (int j = 0; j < kMbVpt; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < sg.end) ? u * W : -1;
                if (e[j] >= 0) {
                    fetch_pred<W>(pred_x + out_off + e[j], px[j], sx[j]);
                    fetch_pred<W>(pred_a + out_off + e[j], pa[j], sa[j]);
                    fetch_raw<T, W>(noise + noise_off + e[j], rn[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < kMbVpt; ++j) {
                if (e[j] < 0) continue;
                float x[W], a[W], n[W];
                decode_pred<W>(px[j], sx[j], x);
                decode_pred<W>(pa[j], sa[j], a);
                decode_raw<T, W>(rn[j], n);
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const float dx = __fsub_rn(x[q], n[q]), da = __fsub_rn(a[q], n[q]);
                    acc[0] = fmaf(dx, dx, acc[0]);
                    acc[1] = fmaf(da, da, acc[1]);
                }
            }
        }
        double tot[2];
        if (row_reduce<2>(acc, tot, s, ws, row, sg.begin > 0, red, &flag) && threadIdx.x == 0) {
            sum_x[row] = (float)tot[0];
            sum_a[row] = (float)tot[1];
        }
    }
}

template <typename T>
static int launch_membership_add_noise(const void* x0, const void* a0, const void* noise, const float* ac, int t,
                                       void* xt_x, void* xt_a, long long row0, long long rows, long long n_noise,
                                       long long D, cudaStream_t st) {
    constexpr int N = VecTraits<T>::N;
    const bool vec = (D % N == 0) && aligned16(x0) && aligned16(a0) && aligned16(noise) && aligned16(xt_x) && aligned16(xt_a);
    if (vec) {
        RowSched s = make_row_sched(rows, D, N, kMbOcc);
        membership_add_noise_kernel<T, N><<<s.grid, kThreads, 0, st>>>((const T*)x0, (const T*)a0, (const T*)noise, ac, t,
                                                                       (T*)xt_x, (T*)xt_a, row0, n_noise, s);
    } else {
        RowSched s = make_row_sched(rows, D, 1, kMbOcc);
        membership_add_noise_kernel<T, 1><<<s.grid, kThreads, 0, st>>>((const T*)x0, (const T*)a0, (const T*)noise, ac, t,
                                                                       (T*)xt_x, (T*)xt_a, row0, n_noise, s);
    }
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_membership_sqerr(const float* pred_x, const float* pred_a, const void* noise, float* sum_x,
                                   float* sum_a, void* workspace, long long row0, long long rows, long long n_noise,
                                   long long D, cudaStream_t st) {
    constexpr int N = VecTraits<T>::N;
    const bool vec = (D % N == 0) && aligned16(pred_x) && aligned16(pred_a) && aligned16(noise);
    const RowWorkspace ws = carve_row_workspace(workspace, rows);
    if (vec) {
        RowSched s = make_row_sched(rows, D, N, kMbOcc);
        membership_sqerr_kernel<T, N><<<s.grid, kThreads, 0, st>>>(pred_x, pred_a, (const T*)noise, sum_x, sum_a, row0,
                                                                   n_noise, ws, s);
    } else {
        RowSched s = make_row_sched(rows, D, 1, kMbOcc);
        membership_sqerr_kernel<T, 1><<<s.grid, kThreads, 0, st>>>(pred_x, pred_a, (const T*)noise, sum_x, sum_a, row0,
                                                                   n_noise, ws, s);
    }
    return (int)cudaGetLastError();
}

}  // namespace siss

using namespace siss;

extern "C" {

int siss_membership_add_noise(const void* x0, const void* a0, const void* noise, const float* alphas_cumprod,
                              int T_steps, int64_t timestep, void* xt_x, void* xt_a, int64_t row0, int64_t rows,
                              int64_t n_noise, int64_t D, int dtype, siss_stream_t stream) {
    if (!x0 || !a0 || !noise || !alphas_cumprod || !xt_x || !xt_a || T_steps < 1 || timestep < 0 ||
        timestep >= T_steps || row0 < 0 || rows < 0 || n_noise < 1 || D < 1)
        return SISS_EINVAL;
    if (rows == 0) return SISS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case SISS_F32:  return launch_membership_add_noise<float>(x0, a0, noise, alphas_cumprod, (int)timestep, xt_x, xt_a, row0, rows, n_noise, D, st);
        case SISS_BF16: return launch_membership_add_noise<__nv_bfloat16>(x0, a0, noise, alphas_cumprod, (int)timestep, xt_x, xt_a, row0, rows, n_noise, D, st);
        case SISS_F16:  return launch_membership_add_noise<__half>(x0, a0, noise, alphas_cumprod, (int)timestep, xt_x, xt_a, row0, rows, n_noise, D, st);
        default: return SISS_EUNSUPPORTED;
    }
}

int siss_membership_sqerr(const float* pred_x, const float* pred_a, const void* noise, int dtype, float* sum_x,
                          float* sum_a, void* workspace, int64_t row0, int64_t rows, int64_t n_noise, int64_t D,
                          siss_stream_t stream) {
    if (!pred_x || !pred_a || !noise || !sum_x || !sum_a || !workspace || row0 < 0 || rows < 0 || n_noise < 1 || D < 1)
        return SISS_EINVAL;
    if (rows == 0) return SISS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (dtype) {
        case SISS_F32:  return launch_membership_sqerr<float>(pred_x, pred_a, noise, sum_x, sum_a, workspace, row0, rows, n_noise, D, st);
        case SISS_BF16: return launch_membership_sqerr<__nv_bfloat16>(pred_x, pred_a, noise, sum_x, sum_a, workspace, row0, rows, n_noise, D, st);
        case SISS_F16:  return launch_membership_sqerr<__half>(pred_x, pred_a, noise, sum_x, sum_a, workspace, row0, rows, n_noise, D, st);
        default: return SISS_EUNSUPPORTED;
    }
}

}  // extern "C"
