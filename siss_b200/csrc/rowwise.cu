// K1 (DDPM add_noise), K2 (defensive mixture + importance weights) and the fused K1 o K2.
// See include/siss_b200.h for the reference call sites each entry point replaces.
//
// HBM traffic (s = sizeof(T)):  K1 single 3s, K1 pair 5s, K2 4s, K1oK2 4s bytes per element.
// All three are pure streaming kernels: 128-bit coalesced loads issued in a batch per iteration,
// fp32 math in registers, 128-bit stores. K2 additionally reduces three sums per row with
// warp shuffles -> smem -> fixed-order cross-CTA combine.

#include "bulkpipe.cuh"
#include "noise.cuh"
#include "philox.cuh"

#include <cstdlib>

namespace siss {

// Current CUDA device, clamped to the size of the per-device state tables (function attributes and SM counts belong
// to a device/context, not to the process: one process may drive several GPUs).
int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < kMaxDevices ? dev : kMaxDevices - 1;
}

int cached_sm_count() {
    static int sms[kMaxDevices] = {0};
    const int slot = current_device_slot();
    if (sms[slot] == 0) {
        int v = 0, dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = kNumSMsB200;
        sms[slot] = v;
    }
    return sms[slot];
}

int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    if (!e || !*e) return dflt;
    const int v = std::atoi(e);
    return v >= 0 ? v : dflt;
}

// A/B switch for measurements: SISS_NO_TMA=1 forces the plain-LDG kernels on the vector path too.
bool use_tma_pipeline() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("SISS_NO_TMA");
        v = (e && e[0] == '1') ? 0 : 1;
    }
    return v == 1;
}

// ---------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------
constexpr int kK1Vpt = 2;
constexpr int kK1Occ = 4;

template <typename T, int W, int NSRC>
__global__ void __launch_bounds__(kThreads, kK1Occ)
add_noise_kernel(const T* __restrict__ x0, const T* __restrict__ a0, const T* __restrict__ noise,
                 const int64_t* __restrict__ ts, const float* __restrict__ ac, int T_steps,
                 T* __restrict__ xt_x, T* __restrict__ xt_a, RowSched s) {
    constexpr int VPT = kK1Vpt;
    long long u0, u1;
    cta_span(s, u0, u1);
    if (u0 >= u1) return;
    for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
        const RowSeg sg = row_segment(s, u0, u1, row);
        float sa, s1;
        noise_coeffs<T>(ac, wrap_timestep(ts[row], T_steps), sa, s1);
        const long long rowoff = row * s.D;
        for (long long ub = sg.begin; ub < sg.end; ub += (long long)kThreads * VPT) {
            RawUnit<T, W> rx[VPT], ra[VPT], rn[VPT];
            long long e[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < sg.end) ? (rowoff + u * W) : -1;
                if (e[j] >= 0) {
                    fetch_raw<T, W>(x0 + e[j], rx[j]);
                    if (NSRC == 2) fetch_raw<T, W>(a0 + e[j], ra[j]);
                    fetch_raw<T, W>(noise + e[j], rn[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (e[j] < 0) continue;
                float x[W], n[W], o[W];
                decode_raw<T, W>(rx[j], x);
                decode_raw<T, W>(rn[j], n);
#pragma unroll
                for (int k = 0; k < W; ++k) o[k] = noised<T>(sa, s1, x[k], n[k]);
                store_unit<T, W>(xt_x + e[j], o);
                if (NSRC == 2) {
                    decode_raw<T, W>(ra[j], x);
#pragma unroll
                    for (int k = 0; k < W; ++k) o[k] = noised<T>(sa, s1, x[k], n[k]);
                    store_unit<T, W>(xt_a + e[j], o);
                }
            }
        }
    }
}

// TMA-pipelined K1 (vector path): see bulkpipe.cuh.
template <typename T, int NSRC>
struct AddNoiseOp {
    static constexpr int W = VecTraits<T>::N;
    static constexpr int NIN = NSRC + 1;
    static constexpr int K = 0;
    static constexpr int kOcc = 3;
    static constexpr int kStages = 6;
    __host__ __device__ static constexpr int ub(int) { return 16; }
    struct Params {
        const T* x0; const T* a0; const T* noise; const int64_t* ts; const float* ac; int T_steps;
        T* xt_x; T* xt_a;
    };
    struct Row { float sa, s1; };
    __device__ static __forceinline__ Row row_begin(const Params& p, long long row) {
        Row r;
        noise_coeffs<T>(p.ac, wrap_timestep(p.ts[row], p.T_steps), r.sa, r.s1);
        return r;
    }
    __device__ static __forceinline__ const char* stream(const Params& p, long long, int i) {
        if (i == 0) return reinterpret_cast<const char*>(p.noise);
        if (i == 1) return reinterpret_cast<const char*>(p.x0);
        return reinterpret_cast<const char*>(p.a0);
    }
    __device__ static __forceinline__ void unit(const Params& p, const Row& r, const uint4 (&in)[NIN][2],
                                                long long unit_index, float (&)[1]) {
        stg_stream(p.xt_x + unit_index * W, VecTraits<T>::noised_unit(r.sa, r.s1, in[1][0], in[0][0]));
        if constexpr (NSRC == 2)
            stg_stream(p.xt_a + unit_index * W, VecTraits<T>::noised_unit(r.sa, r.s1, in[2][0], in[0][0]));
    }
    __device__ static __forceinline__ void row_end(const Params&, const Row&, long long, const double (&)[1]) {}
};

template <typename T, int NSRC>
static int launch_add_noise(const void* x0, const void* a0, const void* noise, const int64_t* ts,
                            const float* ac, int T_steps, void* xt_x, void* xt_a,
                            long long B, long long D, cudaStream_t st) {
    constexpr int N = VecTraits<T>::N;
    bool vec = (D % N == 0) && aligned16(x0) && aligned16(noise) && aligned16(xt_x);
    if (NSRC == 2) vec = vec && aligned16(a0) && aligned16(xt_a);
    if (vec && use_tma_pipeline()) {
        using Op = AddNoiseOp<T, NSRC>;
        typename Op::Params p{(const T*)x0, (const T*)a0, (const T*)noise, ts, ac, T_steps, (T*)xt_x, (T*)xt_a};
        return launch_pipe<Op>(p, RowWorkspace{nullptr, nullptr}, B, D, N, st);
    }
    if (vec) {
        RowSched s = make_row_sched(B, D, N, kK1Occ);
        add_noise_kernel<T, N, NSRC><<<s.grid, kThreads, 0, st>>>(
            (const T*)x0, (const T*)a0, (const T*)noise, ts, ac, T_steps, (T*)xt_x, (T*)xt_a, s);
    } else {
        RowSched s = make_row_sched(B, D, 1, kK1Occ);
        add_noise_kernel<T, 1, NSRC><<<s.grid, kThreads, 0, st>>>(
            (const T*)x0, (const T*)a0, (const T*)noise, ts, ac, T_steps, (T*)xt_x, (T*)xt_a, s);
    }
    return (int)cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// K2 and K1 o K2
// ---------------------------------------------------------------------------------------------
constexpr int kK2Vpt = 2;
constexpr int kK2Occ = 3;

// Row epilogue (one thread): losses/ddpm_deletion_loss.py:34,38 (division by 2 sigma^2) and
// :41-45 (ratios and weights), in the reference's fp32 op order. `sd` is the directly
// accumulated sum of (r_x^2 - r_a^2), so delta = dist_x - dist_a without the cancellation.
__device__ __forceinline__ void finalize_row_weights(double sx, double sa, double sd, float sigma,
                                                     float lam, float one_m_lam, long long row,
                                                     float* __restrict__ dist_x, float* __restrict__ dist_a,
                                                     float* __restrict__ w_x, float* __restrict__ w_a) {
    const float two_s2 = __fmul_rn(2.0f, __fmul_rn(sigma, sigma));
    const float dx = __fdiv_rn((float)sx, two_s2);
    const float da = __fdiv_rn((float)sa, two_s2);
    const float delta = __fdiv_rn((float)sd, two_s2);
    const float r_ax = expf(delta);    // ratio_a_x = exp(dist_x - dist_a)
    const float r_xa = expf(-delta);   // ratio_x_a = exp(dist_a - dist_x)
    dist_x[row] = dx;
    dist_a[row] = da;
    w_x[row] = __fdiv_rn(1.0f, __fadd_rn(one_m_lam, __fmul_rn(lam, r_ax)));
    w_a[row] = __fdiv_rn(1.0f, __fadd_rn(__fmul_rn(one_m_lam, r_xa), lam));
}

template <typename T, int W, bool FUSED_NOISE, bool RNG>
__global__ void __launch_bounds__(kThreads, kK2Occ)
mixture_kernel(const T* __restrict__ src_x,   // !FUSED: noisy keep batch      FUSED: unused
               const T* __restrict__ src_a,   // !FUSED: noisy forget batch    FUSED: unused
               const T* __restrict__ x0, const T* __restrict__ a0,
               const T* __restrict__ noise,   // FUSED only
               const uint8_t* __restrict__ keep, const int64_t* __restrict__ ts,
               const float* __restrict__ ac,  // FUSED only
               const float* __restrict__ gamma, const float* __restrict__ sigma, int T_steps,
               float lam, float one_m_lam,
               T* __restrict__ x_mix, float* __restrict__ dist_x, float* __restrict__ dist_a,
               float* __restrict__ w_x, float* __restrict__ w_a,
               RngStream rng, const unsigned long long* __restrict__ d_draw, unsigned long long elem_offset,
               T* __restrict__ noise_out,   // RNG only
               RowWorkspace ws, RowSched s) {
    if (RNG) rng_draw_from_device(rng, d_draw);
    static_assert(!RNG || FUSED_NOISE, "in-kernel noise only makes sense fused with add_noise");
    constexpr int VPT = kK2Vpt;
    __shared__ float red[3 * kWarps];
    __shared__ int flag;
    long long u0, u1;
    cta_span(s, u0, u1);
    if (u0 >= u1) return;

    for (long long row = u0 / s.upr; row * s.upr < u1; ++row) {
        const RowSeg sg = row_segment(s, u0, u1, row);
        const int t = wrap_timestep(ts[row], T_steps);
        const bool k = keep[row] != 0;
        const float g = gamma[t];
        float sa = 0.f, s1 = 0.f;
        if (FUSED_NOISE) noise_coeffs<T>(ac, t, sa, s1);
        // third stream: eps when fused, otherwise the selected noisy row
        const T* __restrict__ third = FUSED_NOISE ? noise : (k ? src_x : src_a);
        const long long rowoff = row * s.D;

        float acc[3] = {0.f, 0.f, 0.f};  // sum r_x^2, sum r_a^2, sum (r_x^2 - r_a^2)
        for (long long ub = sg.begin; ub < sg.end; ub += (long long)kThreads * VPT) {
            RawUnit<T, W> rx[VPT], ra[VPT], rs[VPT];
            long long e[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < sg.end) ? (rowoff + u * W) : -1;
                if (e[j] >= 0) {
                    fetch_raw<T, W>(x0 + e[j], rx[j]);
                    fetch_raw<T, W>(a0 + e[j], ra[j]);
                    if (!RNG) fetch_raw<T, W>(third + e[j], rs[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (e[j] < 0) continue;
                float x[W], a[W], v[W], m[W];
                decode_raw<T, W>(rx[j], x);
                decode_raw<T, W>(ra[j], a);
                if constexpr (RNG) {
                    // eps of this unit from the counter-based stream, rounded to the latent dtype exactly as a
                    // materialised noise tensor would hold it (siss_randn), optionally written out
                    rng_normals<W>(rng, elem_offset + (unsigned long long)e[j], v);
#pragma unroll
                    for (int q = 0; q < W; ++q) v[q] = VecTraits<T>::round(v[q]);
                    if (noise_out) store_unit<T, W>(noise_out + e[j], v);
                } else {
                    decode_raw<T, W>(rs[j], v);
                }
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    m[q] = FUSED_NOISE ? noised<T>(sa, s1, k ? x[q] : a[q], v[q]) : v[q];
                    const float r_x = __fsub_rn(m[q], __fmul_rn(g, x[q]));
                    const float r_a = __fsub_rn(m[q], __fmul_rn(g, a[q]));
                    acc[0] = fmaf(r_x, r_x, acc[0]);
                    acc[1] = fmaf(r_a, r_a, acc[1]);
                    acc[2] = fmaf(r_x - r_a, r_x + r_a, acc[2]);
                }
                store_unit<T, W>(x_mix + e[j], m);
            }
        }

        double tot[3];
        if (row_reduce<3>(acc, tot, s, ws, row, sg.begin > 0, red, &flag) && threadIdx.x == 0)
            finalize_row_weights(tot[0], tot[1], tot[2], sigma[t], lam, one_m_lam, row, dist_x, dist_a, w_x, w_a);
    }
}

// TMA-pipelined K2 / K1oK2 (vector path): see bulkpipe.cuh.
template <typename T, bool FUSED_NOISE, bool RNG = false, int OCC = 3, int STAGES = 6>
struct MixtureOp {
    static_assert(!RNG || FUSED_NOISE, "in-kernel noise only makes sense fused with add_noise");
    static constexpr int W = VecTraits<T>::N;
    static constexpr int NIN = RNG ? 2 : 3;   // RNG: eps is generated in registers, only x0 and a0 are streamed
    static constexpr int K = 3;
    static constexpr int kOcc = OCC;
    static constexpr int kStages = STAGES;   // default 6 x 12 KB = 72 KB per CTA, 3 CTAs/SM (SISS_K12_VARIANT: other shapes, A/B)
    __host__ __device__ static constexpr int ub(int) { return 16; }
    struct Params {
        const T* src_x; const T* src_a; const T* x0; const T* a0; const T* noise;
        const uint8_t* keep; const int64_t* ts; const float* ac; const float* gamma; const float* sigma;
        int T_steps; float lam, one_m_lam;
        T* x_mix; float* dist_x; float* dist_a; float* w_x; float* w_a;
        RngStream rng; const unsigned long long* d_draw; unsigned long long elem_offset; T* noise_out;   // RNG only
    };
    struct Row { int t; bool k; float g, sa, s1; RngStream rs; };
    __device__ static __forceinline__ Row row_begin(const Params& p, long long row) {
        Row r;
        if constexpr (RNG) { r.rs = p.rng; rng_draw_from_device(r.rs, p.d_draw); }   // draw index may live on the device
        r.t = wrap_timestep(p.ts[row], p.T_steps);
        r.k = p.keep[row] != 0;
        r.g = p.gamma[r.t];
        r.sa = 0.f; r.s1 = 0.f;
        if (FUSED_NOISE) noise_coeffs<T>(p.ac, r.t, r.sa, r.s1);
        return r;
    }
    __device__ static __forceinline__ const char* stream(const Params& p, long long row, int i) {
        if (i == 0) return reinterpret_cast<const char*>(p.x0);
        if (i == 1) return reinterpret_cast<const char*>(p.a0);
        return reinterpret_cast<const char*>(FUSED_NOISE ? p.noise : (p.keep[row] != 0 ? p.src_x : p.src_a));
    }
    __device__ static __forceinline__ void unit(const Params& p, const Row& r, const uint4 (&in)[NIN][2],
                                                long long unit_index, float (&acc)[3]) {
        float x[W], a[W], m[W];
        VecTraits<T>::unpack(in[0][0], x);
        VecTraits<T>::unpack(in[1][0], a);
        // x_t of the selected source (packed 16-bit arithmetic, no conversions), or the given noisy row
        uint4 third;
        if constexpr (RNG) {
            float z[W];
            rng_normals<W>(r.rs, p.elem_offset + (unsigned long long)unit_index * W, z);
            third = VecTraits<T>::pack(z);     // eps rounded to the latent dtype, as siss_randn stores it
            if (p.noise_out) stg_stream(p.noise_out + unit_index * W, third);
        } else {
            third = in[NIN - 1][0];
        }
        const uint4 mraw = FUSED_NOISE ? VecTraits<T>::noised_unit(r.sa, r.s1, r.k ? in[0][0] : in[1][0], third) : third;
        VecTraits<T>::unpack(mraw, m);
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const float r_x = __fsub_rn(m[q], __fmul_rn(r.g, x[q]));
            const float r_a = __fsub_rn(m[q], __fmul_rn(r.g, a[q]));
            acc[0] = fmaf(r_x, r_x, acc[0]);
            acc[1] = fmaf(r_a, r_a, acc[1]);
            acc[2] = fmaf(r_x - r_a, r_x + r_a, acc[2]);
        }
        stg_stream(p.x_mix + unit_index * W, mraw);
    }
    __device__ static __forceinline__ void row_end(const Params& p, const Row& r, long long row, const double (&tot)[3]) {
        finalize_row_weights(tot[0], tot[1], tot[2], p.sigma[r.t], p.lam, p.one_m_lam, row, p.dist_x, p.dist_a,
                             p.w_x, p.w_a);
    }
};

template <typename T, bool FUSED, bool RNG = false>
static int launch_mixture(const void* src_x, const void* src_a, const void* x0, const void* a0,
                          const void* noise, const uint8_t* keep, const int64_t* ts, const float* ac,
                          const float* gamma, const float* sigma, int T_steps, double lambd,
                          void* x_mix, float* dist_x, float* dist_a, float* w_x, float* w_a,
                          void* workspace, long long B, long long D, cudaStream_t st,
                          RngStream rng = RngStream{0, 0, 0, 0}, const unsigned long long* d_draw = nullptr,
                          unsigned long long elem_offset = 0, void* noise_out = nullptr) {
    constexpr int N = VecTraits<T>::N;
    // torch wraps the python scalars `lambd` and `1 - lambd` to the tensor dtype (fp32).
    const float lam = (float)lambd;
    const float one_m = (float)(1.0 - lambd);
    bool vec = (D % N == 0) && aligned16(x0) && aligned16(a0) && aligned16(x_mix);
    if (RNG) vec = vec && aligned16(noise_out) && (elem_offset % 4 == 0);   // one Philox call = 4 consecutive elements
    else vec = vec && (FUSED ? aligned16(noise) : (aligned16(src_x) && aligned16(src_a)));
    RowWorkspace ws = carve_row_workspace(workspace, B);
    if (vec && use_tma_pipeline()) {
        using Op = MixtureOp<T, FUSED, RNG>;
        typename Op::Params p{(const T*)src_x, (const T*)src_a, (const T*)x0, (const T*)a0, (const T*)noise, keep, ts, ac,
                              gamma, sigma, T_steps, lam, one_m, (T*)x_mix, dist_x, dist_a, w_x, w_a,
                              rng, d_draw, elem_offset, (T*)noise_out};
        if constexpr (FUSED && !RNG) {
            static const int variant = env_int("SISS_K12_VARIANT", 0);    // A/B knob: ring shape
            if (variant == 1) {   // 2 CTAs/SM x 9 stages
                using O1 = MixtureOp<T, FUSED, RNG, 2, 9>;
                typename O1::Params q{p.src_x, p.src_a, p.x0, p.a0, p.noise, p.keep, p.ts, p.ac, p.gamma, p.sigma, p.T_steps,
                                      p.lam, p.one_m_lam, p.x_mix, p.dist_x, p.dist_a, p.w_x, p.w_a, p.rng, p.d_draw,
                                      p.elem_offset, p.noise_out};
                return launch_pipe<O1>(q, ws, B, D, N, st);
            }
            if (variant == 2) {   // 3 CTAs/SM x 4 stages
                using O2 = MixtureOp<T, FUSED, RNG, 3, 4>;
                typename O2::Params q{p.src_x, p.src_a, p.x0, p.a0, p.noise, p.keep, p.ts, p.ac, p.gamma, p.sigma, p.T_steps,
                                      p.lam, p.one_m_lam, p.x_mix, p.dist_x, p.dist_a, p.w_x, p.w_a, p.rng, p.d_draw,
                                      p.elem_offset, p.noise_out};
                return launch_pipe<O2>(q, ws, B, D, N, st);
            }
            if (variant == 3) {   // 1 CTA/SM x 18 stages
                using O3 = MixtureOp<T, FUSED, RNG, 1, 18>;
                typename O3::Params q{p.src_x, p.src_a, p.x0, p.a0, p.noise, p.keep, p.ts, p.ac, p.gamma, p.sigma, p.T_steps,
                                      p.lam, p.one_m_lam, p.x_mix, p.dist_x, p.dist_a, p.w_x, p.w_a, p.rng, p.d_draw,
                                      p.elem_offset, p.noise_out};
                return launch_pipe<O3>(q, ws, B, D, N, st);
            }
        }
        return launch_pipe<Op>(p, ws, B, D, N, st);
    }
    if (vec) {
        RowSched s = make_row_sched(B, D, N, kK2Occ);
        static const int over = env_int("SISS_LDG_OVERSUB", 1);
        oversubscribe(s, over, kK2Vpt);
        mixture_kernel<T, N, FUSED, RNG><<<s.grid, kThreads, 0, st>>>(
            (const T*)src_x, (const T*)src_a, (const T*)x0, (const T*)a0, (const T*)noise, keep, ts, ac,
            gamma, sigma, T_steps, lam, one_m, (T*)x_mix, dist_x, dist_a, w_x, w_a, rng, d_draw, elem_offset, (T*)noise_out, ws, s);
    } else {
        RowSched s = make_row_sched(B, D, 1, kK2Occ);
        mixture_kernel<T, 1, FUSED, RNG><<<s.grid, kThreads, 0, st>>>(
            (const T*)src_x, (const T*)src_a, (const T*)x0, (const T*)a0, (const T*)noise, keep, ts, ac,
            gamma, sigma, T_steps, lam, one_m, (T*)x_mix, dist_x, dist_a, w_x, w_a, rng, d_draw, elem_offset, (T*)noise_out, ws, s);
    }
    return (int)cudaGetLastError();
}

}  // namespace siss

using namespace siss;

#define SISS_DISPATCH_DTYPE(dtype, ...)                                          \
    switch (dtype) {                                                             \
        case SISS_F32:  { using T = float;          return __VA_ARGS__; }        \
        case SISS_BF16: { using T = __nv_bfloat16;  return __VA_ARGS__; }        \
        case SISS_F16:  { using T = __half;         return __VA_ARGS__; }        \
        default: return SISS_EUNSUPPORTED;                                       \
    }

extern "C" {

int64_t siss_row_workspace_bytes(int64_t B) {
    if (B < 1) B = 1;
    return row_ws_bytes(B);
}

int siss_add_noise(const void* x0, const void* noise, const int64_t* timesteps,
                   const float* alphas_cumprod, int T_steps, void* xt,
                   int64_t B, int64_t D, int dtype, siss_stream_t stream) {
    if (!x0 || !noise || !timesteps || !alphas_cumprod || !xt || B < 0 || D < 0 || T_steps < 1) return SISS_EINVAL;
    if (B == 0 || D == 0) return SISS_OK;
    SISS_DISPATCH_DTYPE(dtype, (launch_add_noise<T, 1>(x0, nullptr, noise, timesteps, alphas_cumprod, T_steps,
                                                       xt, nullptr, B, D, (cudaStream_t)stream)));
}

int siss_add_noise_pair(const void* x0, const void* a0, const void* noise, const int64_t* timesteps,
                        const float* alphas_cumprod, int T_steps, void* xt_x, void* xt_a,
                        int64_t B, int64_t D, int dtype, siss_stream_t stream) {
    if (!x0 || !a0 || !noise || !timesteps || !alphas_cumprod || !xt_x || !xt_a || B < 0 || D < 0 || T_steps < 1)
        return SISS_EINVAL;
    if (B == 0 || D == 0) return SISS_OK;
    SISS_DISPATCH_DTYPE(dtype, (launch_add_noise<T, 2>(x0, a0, noise, timesteps, alphas_cumprod, T_steps,
                                                       xt_x, xt_a, B, D, (cudaStream_t)stream)));
}

int siss_mixture_weights(const void* xt_x, const void* xt_a, const void* x0, const void* a0,
                         const uint8_t* keep_mask, const int64_t* timesteps,
                         const float* gamma, const float* sigma, int T_steps, double lambd,
                         void* x_mix, float* dist_x, float* dist_a, float* w_x, float* w_a,
                         void* workspace, int64_t B, int64_t D, int dtype, siss_stream_t stream) {
    if (!xt_x || !xt_a || !x0 || !a0 || !keep_mask || !timesteps || !gamma || !sigma || !x_mix || !dist_x ||
        !dist_a || !w_x || !w_a || !workspace || B < 0 || D < 1 || T_steps < 1)
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_DTYPE(dtype, (launch_mixture<T, false>(xt_x, xt_a, x0, a0, nullptr, keep_mask, timesteps, nullptr,
                                                         gamma, sigma, T_steps, lambd, x_mix, dist_x, dist_a, w_x,
                                                         w_a, workspace, B, D, (cudaStream_t)stream)));
}

int siss_add_noise_mixture(const void* x0, const void* a0, const void* noise,
                           const uint8_t* keep_mask, const int64_t* timesteps,
                           const float* alphas_cumprod, const float* gamma, const float* sigma, int T_steps,
                           double lambd,
                           void* x_mix, float* dist_x, float* dist_a, float* w_x, float* w_a,
                           void* workspace, int64_t B, int64_t D, int dtype, siss_stream_t stream) {
    if (!x0 || !a0 || !noise || !keep_mask || !timesteps || !alphas_cumprod || !gamma || !sigma || !x_mix ||
        !dist_x || !dist_a || !w_x || !w_a || !workspace || B < 0 || D < 1 || T_steps < 1)
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_DTYPE(dtype, (launch_mixture<T, true>(nullptr, nullptr, x0, a0, noise, keep_mask, timesteps,
                                                        alphas_cumprod, gamma, sigma, T_steps, lambd, x_mix, dist_x,
                                                        dist_a, w_x, w_a, workspace, B, D, (cudaStream_t)stream)));
}

int siss_add_noise_mixture_rng(const void* x0, const void* a0, const uint8_t* keep_mask, const int64_t* timesteps,
                               const float* alphas_cumprod, const float* gamma, const float* sigma, int T_steps,
                               double lambd, uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
                               void* x_mix, void* noise_out, float* dist_x, float* dist_a, float* w_x, float* w_a,
                               void* workspace, int64_t B, int64_t D, int dtype, siss_stream_t stream) {
    if (!x0 || !a0 || !keep_mask || !timesteps || !alphas_cumprod || !gamma || !sigma || !x_mix ||
        !dist_x || !dist_a || !w_x || !w_a || !workspace || B < 0 || D < 1 || T_steps < 1 || (draw >> 62))
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_DTYPE(dtype, (launch_mixture<T, true, true>(nullptr, nullptr, x0, a0, nullptr, keep_mask, timesteps,
                                                              alphas_cumprod, gamma, sigma, T_steps, lambd, x_mix, dist_x,
                                                              dist_a, w_x, w_a, workspace, B, D, (cudaStream_t)stream,
                                                              make_rng_stream(seed, draw, kRngNoise),
                                                              (const unsigned long long*)d_draw, elem_offset, noise_out)));
}

}  // extern "C"
