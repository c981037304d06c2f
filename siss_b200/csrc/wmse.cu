// K3: weighted epsilon-MSE forward + backward into the UNet output, the API-compatible
// materialising forward / general backward, and the plain squared-error family used by the
// No-IS / EraseDiff / NegGrad / naive losses. Reference call sites: include/siss_b200.h.
//
// fp32 op order follows eager exactly (no FMA contraction on the value path), so per-element
// outputs are bit-identical to the reference's ATen sequence; only the per-row sums differ in
// reduction order.

#include "bulkpipe.cuh"
#include "philox.cuh"

namespace siss {

bool use_tma_pipeline();

// u_x = pred - (x_mix - gamma*x0)/sigma  — losses/ddpm_deletion_loss.py:26,29 in eager's order:
// mul, sub, div, sub, each rounded to fp32.
// The division: sigma is constant per row, so the quotient is formed as (float)((double)r * inv_sg)
// with inv_sg = 1.0 / (double)sigma computed once per row. That IS the correctly rounded fp32 quotient:
// the product carries two fp64 roundings (relative error < 2^-52), while a quotient of two 24-bit
// significands is never closer than ~2^-49 (relative) to an fp32 rounding boundary (the classic
// "2p+2 bits suffice" argument for division), so the final rounding decides exactly as IEEE division
// does. 2 cvt + 1 DMUL instead of the ~9-instruction div.rn routine; K3 is issue-heavy, this matters.
__device__ __forceinline__ float residual(float pred, float xm, float g, double inv_sg, float x) {
    const float r = __fsub_rn(xm, __fmul_rn(g, x));
    return __fsub_rn(pred, (float)((double)r * inv_sg));
}

// ---------------------------------------------------------------------------------------------
// K3 fast path. TP = dtype of pred and of both gradients, T = dtype of x_mix / x0 / a0.
// One unit = W elements where W = elements per 128-bit access of the NARROWER-count stream:
// we take W = VecTraits<T>::N (4 or 8); pred/grad then need W/VecTraits<TP>::N accesses each.
// ---------------------------------------------------------------------------------------------
constexpr int kK3Vpt = 2;
constexpr int kK3Occ = 2;

template <typename TP, int W, bool VEC>
struct PredIO {
    static constexpr int NP = VecTraits<TP>::N;
    static constexpr int PARTS = VEC ? (W / NP > 0 ? W / NP : 1) : 1;
    // raw storage for one unit of pred
    struct Raw { uint4 v[PARTS]; float s; };
    __device__ static __forceinline__ void fetch(const TP* p, Raw& r) {
        if constexpr (!VEC) { r.s = VecTraits<TP>::load1(p); }
        else {
#pragma unroll
            for (int i = 0; i < PARTS; ++i) r.v[i] = ldg_stream(p + i * NP);
        }
    }
    __device__ static __forceinline__ void decode(const Raw& r, float (&f)[W]) {
        if constexpr (!VEC) { f[0] = r.s; }
        else {
#pragma unroll
            for (int i = 0; i < PARTS; ++i) {
                float tmp[NP];
                VecTraits<TP>::unpack(r.v[i], tmp);
#pragma unroll
                for (int q = 0; q < NP; ++q) if (i * NP + q < W) f[i * NP + q] = tmp[q];
            }
        }
    }
    __device__ static __forceinline__ void store(TP* p, const float (&f)[W]) {
        if constexpr (!VEC) { VecTraits<TP>::store1(p, f[0]); }
        else {
#pragma unroll
            for (int i = 0; i < PARTS; ++i) {
                float tmp[NP];
#pragma unroll
                for (int q = 0; q < NP; ++q) tmp[q] = (i * NP + q < W) ? f[i * NP + q] : 0.f;
                stg_stream(p + i * NP, VecTraits<TP>::pack(tmp));
            }
        }
    }
};

template <typename TP, typename T, int W, bool VEC>
__global__ void __launch_bounds__(kThreads, kK3Occ)
wmse_fwd_bwd_kernel(const TP* __restrict__ pred, const T* __restrict__ x_mix, const T* __restrict__ x0,
                    const T* __restrict__ a0, const int64_t* __restrict__ ts,
                    const float* __restrict__ gamma, const float* __restrict__ sigma, int T_steps,
                    const float* __restrict__ w_x, const float* __restrict__ w_a, float go_x, float go_a,
                    TP* __restrict__ grad_x, TP* __restrict__ grad_a,
                    float* __restrict__ row_loss_x, float* __restrict__ row_loss_a,
                    RowWorkspace ws, RowSched rt) {
    constexpr int VPT = kK3Vpt;
    using PIO = PredIO<TP, W, VEC>;
    __shared__ float red[2 * kWarps];
    __shared__ int flag;

    long long u0, u1;
    cta_span(rt, u0, u1);
    if (u0 >= u1) return;
    for (long long row = u0 / rt.upr; row * rt.upr < u1; ++row) {
        const RowSeg seg = row_segment(rt, u0, u1, row);
        const long long rowoff = row * rt.D;
        const int t = wrap_timestep(ts[row], T_steps);
        const float g = gamma[t];
        const double sg = 1.0 / (double)sigma[t];   // reciprocal in fp64, see residual()
        // autograd: grad(weighted_loss) = go ; grad(loss) = go * w  (mul backward, rounded once)
        const float cx = __fmul_rn(go_x, w_x[row]);
        const float ca = __fmul_rn(go_a, w_a[row]);

        float acc[2] = {0.f, 0.f};
        for (long long ub = seg.begin; ub < seg.end; ub += (long long)kThreads * VPT) {
            typename PIO::Raw rp[VPT];
            RawUnit<T, W> rm[VPT], rx[VPT], ra[VPT];
            long long e[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < seg.end) ? (rowoff + u * W) : -1;
                if (e[j] >= 0) {
                    PIO::fetch(pred + e[j], rp[j]);
                    fetch_raw<T, W>(x_mix + e[j], rm[j]);
                    fetch_raw<T, W>(x0 + e[j], rx[j]);
                    fetch_raw<T, W>(a0 + e[j], ra[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (e[j] < 0) continue;
                float p[W], m[W], x[W], a[W], gx[W], ga[W];
                PIO::decode(rp[j], p);
                decode_raw<T, W>(rm[j], m);
                decode_raw<T, W>(rx[j], x);
                decode_raw<T, W>(ra[j], a);
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const float ux = residual(p[q], m[q], g, sg, x[q]);
                    const float ua = residual(p[q], m[q], g, sg, a[q]);
                    acc[0] = fmaf(ux, ux, acc[0]);
                    acc[1] = fmaf(ua, ua, acc[1]);
                    // pow backward: grad * (2 * u)
                    gx[q] = __fmul_rn(cx, __fmul_rn(2.0f, ux));
                    ga[q] = __fmul_rn(ca, __fmul_rn(2.0f, ua));
                }
                PIO::store(grad_x + e[j], gx);
                PIO::store(grad_a + e[j], ga);
            }
        }
        double tot[2];
        if (row_reduce<2>(acc, tot, rt, ws, row, seg.begin > 0, red, &flag) && threadIdx.x == 0) {
            row_loss_x[row] = (float)tot[0];
            row_loss_a[row] = (float)tot[1];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// API-compatible forward: materialise loss_x, loss_a, w_x*loss_x, w_a*loss_a (all fp32).
// ---------------------------------------------------------------------------------------------
template <typename TP, typename T, int W, bool VEC>
__global__ void __launch_bounds__(kThreads, kK3Occ)
wmse_fwd_kernel(const TP* __restrict__ pred, const T* __restrict__ x_mix, const T* __restrict__ x0,
                const T* __restrict__ a0, const int64_t* __restrict__ ts,
                const float* __restrict__ gamma, const float* __restrict__ sigma, int T_steps,
                const float* __restrict__ w_x, const float* __restrict__ w_a,
                float* __restrict__ loss_x, float* __restrict__ loss_a,
                float* __restrict__ wloss_x, float* __restrict__ wloss_a, RowSched rt) {
    constexpr int VPT = kK3Vpt;
    using PIO = PredIO<TP, W, VEC>;
    using OIO = PredIO<float, W, VEC>;
    long long u0, u1;
    cta_span(rt, u0, u1);
    if (u0 >= u1) return;
    for (long long row = u0 / rt.upr; row * rt.upr < u1; ++row) {
        const RowSeg seg = row_segment(rt, u0, u1, row);
        const long long rowoff = row * rt.D;
        const int t = wrap_timestep(ts[row], T_steps);
        const float g = gamma[t];
        const double sg = 1.0 / (double)sigma[t];   // reciprocal in fp64, see residual()
        const float wx = w_x[row], wa = w_a[row];
        for (long long ub = seg.begin; ub < seg.end; ub += (long long)kThreads * VPT) {
            typename PIO::Raw rp[VPT];
            RawUnit<T, W> rm[VPT], rx[VPT], ra[VPT];
            long long e[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < seg.end) ? (rowoff + u * W) : -1;
                if (e[j] >= 0) {
                    PIO::fetch(pred + e[j], rp[j]);
                    fetch_raw<T, W>(x_mix + e[j], rm[j]);
                    fetch_raw<T, W>(x0 + e[j], rx[j]);
                    fetch_raw<T, W>(a0 + e[j], ra[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (e[j] < 0) continue;
                float p[W], m[W], x[W], a[W], lx[W], la[W], o[W];
                PIO::decode(rp[j], p);
                decode_raw<T, W>(rm[j], m);
                decode_raw<T, W>(rx[j], x);
                decode_raw<T, W>(ra[j], a);
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const float ux = residual(p[q], m[q], g, sg, x[q]);
                    const float ua = residual(p[q], m[q], g, sg, a[q]);
                    lx[q] = __fmul_rn(ux, ux);
                    la[q] = __fmul_rn(ua, ua);
                }
                if (loss_x) OIO::store(loss_x + e[j], lx);
                if (loss_a) OIO::store(loss_a + e[j], la);
                if (wloss_x) {
#pragma unroll
                    for (int q = 0; q < W; ++q) o[q] = __fmul_rn(wx, lx[q]);
                    OIO::store(wloss_x + e[j], o);
                }
                if (wloss_a) {
#pragma unroll
                    for (int q = 0; q < W; ++q) o[q] = __fmul_rn(wa, la[q]);
                    OIO::store(wloss_a + e[j], o);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// API-compatible backward. Upstream gradients: absent / broadcast scalar / dense fp32.
// ---------------------------------------------------------------------------------------------
struct GradOut {
    const float* p;
    int stride;  // 0 = broadcast scalar, 1 = dense
};

template <typename TP, typename T, int W, bool VEC>
__global__ void __launch_bounds__(kThreads, kK3Occ)
wmse_bwd_kernel(const TP* __restrict__ pred, const T* __restrict__ x_mix, const T* __restrict__ x0,
                const T* __restrict__ a0, const int64_t* __restrict__ ts,
                const float* __restrict__ gamma, const float* __restrict__ sigma, int T_steps,
                const float* __restrict__ w_x, const float* __restrict__ w_a,
                GradOut go_lx, GradOut go_la, GradOut go_wx, GradOut go_wa,
                TP* __restrict__ grad_pred, RowSched rt) {
    constexpr int VPT = 1;
    using PIO = PredIO<TP, W, VEC>;
    using GIO = PredIO<float, W, VEC>;
    // broadcast upstream scalars (what `.sum()` backward provides)
    const float s_lx = (go_lx.p && go_lx.stride == 0) ? go_lx.p[0] : 0.f;
    const float s_la = (go_la.p && go_la.stride == 0) ? go_la.p[0] : 0.f;
    const float s_wx = (go_wx.p && go_wx.stride == 0) ? go_wx.p[0] : 0.f;
    const float s_wa = (go_wa.p && go_wa.stride == 0) ? go_wa.p[0] : 0.f;
    const bool d_lx = go_lx.p && go_lx.stride != 0, d_la = go_la.p && go_la.stride != 0;
    const bool d_wx = go_wx.p && go_wx.stride != 0, d_wa = go_wa.p && go_wa.stride != 0;
    const bool use_x = go_lx.p || go_wx.p, use_a = go_la.p || go_wa.p;

    long long u0, u1;
    cta_span(rt, u0, u1);
    if (u0 >= u1) return;
    for (long long row = u0 / rt.upr; row * rt.upr < u1; ++row) {
        const RowSeg seg = row_segment(rt, u0, u1, row);
        const long long rowoff = row * rt.D;
        const int t = wrap_timestep(ts[row], T_steps);
        const float g = gamma[t];
        const double sg = 1.0 / (double)sigma[t];   // reciprocal in fp64, see residual()
        const float wx = w_x[row], wa = w_a[row];
        for (long long ub = seg.begin; ub < seg.end; ub += (long long)kThreads * VPT) {
            const long long u = ub + threadIdx.x;
            if (u >= seg.end) continue;
            const long long e = rowoff + u * W;
            typename PIO::Raw rp;
            RawUnit<T, W> rm, rx, ra;
            typename GIO::Raw r1, r2, r3, r4;
            PIO::fetch(pred + e, rp);
            fetch_raw<T, W>(x_mix + e, rm);
            if (use_x) fetch_raw<T, W>(x0 + e, rx);
            if (use_a) fetch_raw<T, W>(a0 + e, ra);
            if (d_lx) GIO::fetch(go_lx.p + e, r1);
            if (d_la) GIO::fetch(go_la.p + e, r2);
            if (d_wx) GIO::fetch(go_wx.p + e, r3);
            if (d_wa) GIO::fetch(go_wa.p + e, r4);
            float p[W], m[W], x[W], a[W], out[W], glx[W], gla[W], gwx[W], gwa[W];
            PIO::decode(rp, p);
            decode_raw<T, W>(rm, m);
            if (use_x) decode_raw<T, W>(rx, x);
            if (use_a) decode_raw<T, W>(ra, a);
            if (d_lx) GIO::decode(r1, glx);
            if (d_la) GIO::decode(r2, gla);
            if (d_wx) GIO::decode(r3, gwx);
            if (d_wa) GIO::decode(r4, gwa);
#pragma unroll
            for (int q = 0; q < W; ++q) {
                float acc = 0.f;
                bool first = true;
                if (use_x) {
                    const float two_u = __fmul_rn(2.0f, residual(p[q], m[q], g, sg, x[q]));
                    if (go_wx.p) {
                        const float c = __fmul_rn(d_wx ? gwx[q] : s_wx, wx);
                        acc = __fmul_rn(c, two_u); first = false;
                    }
                    if (go_lx.p) {
                        const float v = __fmul_rn(d_lx ? glx[q] : s_lx, two_u);
                        acc = first ? v : __fadd_rn(acc, v); first = false;
                    }
                }
                if (use_a) {
                    const float two_u = __fmul_rn(2.0f, residual(p[q], m[q], g, sg, a[q]));
                    if (go_wa.p) {
                        const float c = __fmul_rn(d_wa ? gwa[q] : s_wa, wa);
                        const float v = __fmul_rn(c, two_u);
                        acc = first ? v : __fadd_rn(acc, v); first = false;
                    }
                    if (go_la.p) {
                        const float v = __fmul_rn(d_la ? gla[q] : s_la, two_u);
                        acc = first ? v : __fadd_rn(acc, v); first = false;
                    }
                }
                out[q] = acc;
            }
            PIO::store(grad_pred + e, out);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Plain squared error (flat, n elements). TO = promoted output dtype.
// ---------------------------------------------------------------------------------------------
template <typename TP, typename TT, typename TO>
__global__ void __launch_bounds__(kThreads)
sqerr_fwd_kernel(const TP* __restrict__ pred, const TT* __restrict__ tgt, TO* __restrict__ loss,
                 TO* __restrict__ scaled, float alpha, long long n) {
    using VO = VecTraits<TO>;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float u = VO::round(__fsub_rn(VecTraits<TP>::load1(pred + i), VecTraits<TT>::load1(tgt + i)));
        const float l = VO::round(__fmul_rn(u, u));
        VO::store1(loss + i, l);
        if (scaled) VO::store1(scaled + i, VO::round(__fmul_rn(alpha, l)));  // scalar stays fp32 (opmath)
    }
}

// vectorised variant: all three dtypes fp32 (the autocast case) — 4 elements / thread / access
__global__ void __launch_bounds__(kThreads)
sqerr_fwd_f32v4_kernel(const float* __restrict__ pred, const float* __restrict__ tgt, float* __restrict__ loss,
                       float* __restrict__ scaled, float alpha, long long nvec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        float p[4], t[4], l[4], s[4];
        VecTraits<float>::unpack(ldg_stream(pred + 4 * i), p);
        VecTraits<float>::unpack(ldg_stream(tgt + 4 * i), t);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float u = __fsub_rn(p[q], t[q]);
            l[q] = __fmul_rn(u, u);
            s[q] = __fmul_rn(alpha, l[q]);
        }
        stg_stream(loss + 4 * i, VecTraits<float>::pack(l));
        if (scaled) stg_stream(scaled + 4 * i, VecTraits<float>::pack(s));
    }
}

template <typename TP, typename TT, typename TO>
__global__ void __launch_bounds__(kThreads)
sqerr_bwd_kernel(const TP* __restrict__ pred, const TT* __restrict__ tgt,
                 const TO* __restrict__ go_loss, int go_loss_stride,
                 const TO* __restrict__ go_scaled, int go_scaled_stride, float alpha,
                 TP* __restrict__ grad_pred, long long n) {
    using VO = VecTraits<TO>;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float u = VO::round(__fsub_rn(VecTraits<TP>::load1(pred + i), VecTraits<TT>::load1(tgt + i)));
        const float two_u = VO::round(__fmul_rn(2.0f, u));
        float acc = 0.f;
        bool first = true;
        if (go_scaled) {
            const float gs = VO::load1(go_scaled + (go_scaled_stride ? i : 0));
            const float c = VO::round(__fmul_rn(gs, alpha));  // mul backward: grad * alpha (fp32 opmath scalar)
            acc = VO::round(__fmul_rn(c, two_u)); first = false;
        }
        if (go_loss) {
            const float gl = VO::load1(go_loss + (go_loss_stride ? i : 0));
            const float v = VO::round(__fmul_rn(gl, two_u));
            acc = first ? v : VO::round(__fadd_rn(acc, v));
        }
        VecTraits<TP>::store1(grad_pred + i, acc);  // sub backward; cast to pred's dtype
    }
}

// ---------------------------------------------------------------------------------------------
// Dual MSE fast path (No-IS / EraseDiff): two preds, one or two targets, both grads + row sums.
// ---------------------------------------------------------------------------------------------
// RNG_A (opt-in device RNG, EraseDiff): the forget target is not read but drawn in registers — uniform [0, 1) from the
// aux domain of the counter-based stream, rounded to the prediction dtype like torch.rand_like(eps_hat_a)
// (losses/ddpm_deletion_loss.py:75) — and optionally written to tgt_a_out.
template <typename TP, typename TT, int W, bool VEC, bool RNG_A = false>
__global__ void __launch_bounds__(kThreads, kK3Occ)
dual_mse_kernel(const TP* __restrict__ pred_x, const TP* __restrict__ pred_a,
                const TT* __restrict__ tgt_x, const TT* __restrict__ tgt_a, float go_x, float go_a,
                TP* __restrict__ grad_x, TP* __restrict__ grad_a,
                float* __restrict__ row_loss_x, float* __restrict__ row_loss_a,
                RowWorkspace ws, RowSched rt,
                RngStream rng = RngStream{0, 0, 0, 0}, const unsigned long long* __restrict__ d_draw = nullptr,
                unsigned long long elem_offset = 0, TP* __restrict__ tgt_a_out = nullptr) {
    constexpr int VPT = kK3Vpt;
    using PIO = PredIO<TP, W, VEC>;
    using TIO = PredIO<TT, W, VEC>;
    __shared__ float red[2 * kWarps];
    __shared__ int flag;
    const bool shared_tgt = !RNG_A && (tgt_a == tgt_x);
    if (RNG_A) rng_draw_from_device(rng, d_draw);

    long long u0, u1;
    cta_span(rt, u0, u1);
    if (u0 >= u1) return;
    for (long long row = u0 / rt.upr; row * rt.upr < u1; ++row) {
        const RowSeg seg = row_segment(rt, u0, u1, row);
        const long long rowoff = row * rt.D;
        float acc[2] = {0.f, 0.f};
        for (long long ub = seg.begin; ub < seg.end; ub += (long long)kThreads * VPT) {
            typename PIO::Raw rpx[VPT], rpa[VPT];
            typename TIO::Raw rtx[VPT], rta[VPT];
            long long e[VPT];
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const long long u = ub + (long long)j * kThreads + threadIdx.x;
                e[j] = (u < seg.end) ? (rowoff + u * W) : -1;
                if (e[j] >= 0) {
                    PIO::fetch(pred_x + e[j], rpx[j]);
                    PIO::fetch(pred_a + e[j], rpa[j]);
                    TIO::fetch(tgt_x + e[j], rtx[j]);
                    if (!RNG_A && !shared_tgt) TIO::fetch(tgt_a + e[j], rta[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (e[j] < 0) continue;
                float px[W], pa[W], tx[W], ta[W], gx[W], ga[W];
                PIO::decode(rpx[j], px);
                PIO::decode(rpa[j], pa);
                TIO::decode(rtx[j], tx);
                if constexpr (RNG_A) {
                    rng_uniforms<W>(rng, elem_offset + (unsigned long long)e[j], ta);
#pragma unroll
                    for (int q = 0; q < W; ++q) ta[q] = VecTraits<TP>::round(ta[q]);
                    if (tgt_a_out) PIO::store(tgt_a_out + e[j], ta);
                } else if (!shared_tgt) {
                    TIO::decode(rta[j], ta);
                }
#pragma unroll
                for (int q = 0; q < W; ++q) {
                    const float ux = __fsub_rn(px[q], tx[q]);
                    const float ua = __fsub_rn(pa[q], shared_tgt ? tx[q] : ta[q]);
                    acc[0] = fmaf(ux, ux, acc[0]);
                    acc[1] = fmaf(ua, ua, acc[1]);
                    gx[q] = __fmul_rn(go_x, __fmul_rn(2.0f, ux));
                    ga[q] = __fmul_rn(go_a, __fmul_rn(2.0f, ua));
                }
                PIO::store(grad_x + e[j], gx);
                PIO::store(grad_a + e[j], ga);
            }
        }
        double tot[2];
        if (row_reduce<2>(acc, tot, rt, ws, row, seg.begin > 0, red, &flag) && threadIdx.x == 0) {
            row_loss_x[row] = (float)tot[0];
            row_loss_a[row] = (float)tot[1];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// TMA-pipelined fast paths (vector path): see bulkpipe.cuh. A unit is W = 16/sizeof(T) elements of
// the latent dtype; the pred / gradient side of a unit is W * sizeof(TP) bytes = 16 or 32.
// ---------------------------------------------------------------------------------------------
template <typename TP, int W>
__device__ __forceinline__ void decode_pred(const uint4 (&in)[2], float (&f)[W]) {
    constexpr int NP = VecTraits<TP>::N;
    if constexpr (NP >= W) {
        float tmp[NP];
        VecTraits<TP>::unpack(in[0], tmp);
#pragma unroll
        for (int q = 0; q < W; ++q) f[q] = tmp[q];
    } else {
        static_assert(2 * NP == W, "pred unit is at most two 128-bit accesses");
        float a[NP], b[NP];
        VecTraits<TP>::unpack(in[0], a);
        VecTraits<TP>::unpack(in[1], b);
#pragma unroll
        for (int q = 0; q < NP; ++q) { f[q] = a[q]; f[NP + q] = b[q]; }
    }
}

template <typename TP, int W>
__device__ __forceinline__ void store_pred(TP* p, const float (&f)[W]) {
    constexpr int NP = VecTraits<TP>::N;
    static_assert(NP == W || 2 * NP == W, "pred unit is one or two 128-bit accesses");
#pragma unroll
    for (int i = 0; i < W / NP; ++i) {
        float tmp[NP];
#pragma unroll
        for (int q = 0; q < NP; ++q) tmp[q] = f[i * NP + q];
        stg_stream(p + i * NP, VecTraits<TP>::pack(tmp));
    }
}

// OCC / STAGES: CTAs per SM and ring depth (0 = the default below). Other shapes of the same ~200 KB of shared memory
// per SM were measured in round 2 through SISS_K3_VARIANT (tools/rowkernel_ab.py) — see launch_wmse_fwd_bwd.
template <typename TP, typename T, int OCC = 2, int STAGES = 0>
struct WmseFwdBwdOp {
    static constexpr int W = VecTraits<T>::N;
    static constexpr int PB = W * (int)sizeof(TP);   // pred bytes per unit: 16 or 32
    static constexpr int NIN = 4;
    static constexpr int K = 2;
    static constexpr int kOcc = OCC;
    static constexpr int kStages = STAGES ? STAGES : ((PB == 32) ? 5 : 6);   // 5 x 20 KB or 6 x 16 KB per CTA, 2 CTAs/SM
    __host__ __device__ static constexpr int ub(int i) { return i == 0 ? PB : 16; }
    struct Params {
        const TP* pred; const T* x_mix; const T* x0; const T* a0; const int64_t* ts;
        const float* gamma; const float* sigma; int T_steps; const float* w_x; const float* w_a;
        float go_x, go_a; TP* grad_x; TP* grad_a; float* row_loss_x; float* row_loss_a;
    };
    struct Row { float g, cx, ca; double sg; };   // sg = 1 / sigma_t in fp64, see residual()
    __device__ static __forceinline__ Row row_begin(const Params& p, long long row) {
        Row r;
        const int t = wrap_timestep(p.ts[row], p.T_steps);
        r.g = p.gamma[t]; r.sg = 1.0 / (double)p.sigma[t];
        r.cx = __fmul_rn(p.go_x, p.w_x[row]);   // mul backward: grad * w, rounded once
        r.ca = __fmul_rn(p.go_a, p.w_a[row]);
        return r;
    }
    __device__ static __forceinline__ const char* stream(const Params& p, long long, int i) {
        if (i == 0) return reinterpret_cast<const char*>(p.pred);
        if (i == 1) return reinterpret_cast<const char*>(p.x_mix);
        if (i == 2) return reinterpret_cast<const char*>(p.x0);
        return reinterpret_cast<const char*>(p.a0);
    }
    __device__ static __forceinline__ void unit(const Params& p, const Row& r, const uint4 (&in)[NIN][2],
                                                long long unit_index, float (&acc)[2]) {
        float pr[W], m[W], x[W], a[W], gx[W], ga[W];
        decode_pred<TP, W>(in[0], pr);
        VecTraits<T>::unpack(in[1][0], m);
        VecTraits<T>::unpack(in[2][0], x);
        VecTraits<T>::unpack(in[3][0], a);
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const float ux = residual(pr[q], m[q], r.g, r.sg, x[q]);
            const float ua = residual(pr[q], m[q], r.g, r.sg, a[q]);
            acc[0] = fmaf(ux, ux, acc[0]);
            acc[1] = fmaf(ua, ua, acc[1]);
            gx[q] = __fmul_rn(r.cx, __fmul_rn(2.0f, ux));   // pow backward: grad * (2 u)
            ga[q] = __fmul_rn(r.ca, __fmul_rn(2.0f, ua));
        }
        store_pred<TP, W>(p.grad_x + unit_index * W, gx);
        store_pred<TP, W>(p.grad_a + unit_index * W, ga);
    }
    __device__ static __forceinline__ void row_end(const Params& p, const Row&, long long row, const double (&tot)[2]) {
        p.row_loss_x[row] = (float)tot[0];
        p.row_loss_a[row] = (float)tot[1];
    }
};

// Dual MSE: unit = WU elements, WU = max(elements per 128 bits of pred, of target).
template <typename TP, typename TT, bool SHARED_TGT>
struct DualMseOp {
    static constexpr int NP = VecTraits<TP>::N, NT = VecTraits<TT>::N;
    static constexpr int W = NP > NT ? NP : NT;
    static constexpr int PB = W * (int)sizeof(TP);
    static constexpr int TB = W * (int)sizeof(TT);
    static constexpr int NIN = SHARED_TGT ? 3 : 4;
    static constexpr int K = 2;
    static constexpr int kOcc = 2;
    static constexpr int kStages = 4;
    __host__ __device__ static constexpr int ub(int i) { return i < 2 ? PB : TB; }
    struct Params {
        const TP* pred_x; const TP* pred_a; const TT* tgt_x; const TT* tgt_a; float go_x, go_a;
        TP* grad_x; TP* grad_a; float* row_loss_x; float* row_loss_a;
    };
    struct Row { int unused; };
    __device__ static __forceinline__ Row row_begin(const Params&, long long) { return Row{0}; }
    __device__ static __forceinline__ const char* stream(const Params& p, long long, int i) {
        if (i == 0) return reinterpret_cast<const char*>(p.pred_x);
        if (i == 1) return reinterpret_cast<const char*>(p.pred_a);
        if (i == 2) return reinterpret_cast<const char*>(p.tgt_x);
        return reinterpret_cast<const char*>(p.tgt_a);
    }
    __device__ static __forceinline__ void unit(const Params& p, const Row&, const uint4 (&in)[NIN][2],
                                                long long unit_index, float (&acc)[2]) {
        float px[W], pa[W], tx[W], ta[W], gx[W], ga[W];
        decode_pred<TP, W>(in[0], px);
        decode_pred<TP, W>(in[1], pa);
        decode_pred<TT, W>(in[2], tx);
        if constexpr (!SHARED_TGT) decode_pred<TT, W>(in[3], ta);
#pragma unroll
        for (int q = 0; q < W; ++q) {
            const float ux = __fsub_rn(px[q], tx[q]);
            const float ua = __fsub_rn(pa[q], SHARED_TGT ? tx[q] : ta[q]);
            acc[0] = fmaf(ux, ux, acc[0]);
            acc[1] = fmaf(ua, ua, acc[1]);
            gx[q] = __fmul_rn(p.go_x, __fmul_rn(2.0f, ux));
            ga[q] = __fmul_rn(p.go_a, __fmul_rn(2.0f, ua));
        }
        store_pred<TP, W>(p.grad_x + unit_index * W, gx);
        store_pred<TP, W>(p.grad_a + unit_index * W, ga);
    }
    __device__ static __forceinline__ void row_end(const Params& p, const Row&, long long row, const double (&tot)[2]) {
        p.row_loss_x[row] = (float)tot[0];
        p.row_loss_a[row] = (float)tot[1];
    }
};

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
template <typename TP, typename T>
static bool k3_vec_ok(const void* pred, const void* xm, const void* x0, const void* a0, long long D,
                      const void* o1, const void* o2, const void* o3 = nullptr, const void* o4 = nullptr) {
    constexpr int W = VecTraits<T>::N;
    // pred-side accesses are W / NP vectors per unit; need W >= NP (true unless T=fp32 & TP=16-bit)
    if (VecTraits<TP>::N > W) return false;
    bool ok = (D % W == 0) && aligned16(pred) && aligned16(xm) && aligned16(x0) && aligned16(a0);
    ok = ok && aligned16(o1) && aligned16(o2) && aligned16(o3) && aligned16(o4);
    return ok;
}

template <typename TP, typename T>
static int launch_wmse_fwd_bwd(const void* pred, const void* x_mix, const void* x0, const void* a0,
                               const int64_t* ts, const float* gamma, const float* sigma, int T_steps,
                               const float* w_x, const float* w_a, float go_x, float go_a,
                               void* grad_x, void* grad_a, float* row_loss_x, float* row_loss_a,
                               void* workspace, long long B, long long D, cudaStream_t st) {
    constexpr int W = VecTraits<T>::N;
    RowWorkspace ws = carve_row_workspace(workspace, B);
    // Measured on B200 (profiles/r1_microbench_sweep.csv, celeb shape, bf16 latents): the TMA-fed kernel wins
    // while the launch is short (B = 64: 50.7 vs 54.3 us), the register-staged LDG kernel wins once there
    // are many iterations per CTA (B = 256: 166 vs 173 us, B = 1024: 605 vs 654 us). Switch at 4 M units.
    const bool prefer_tma = (B * (D / W)) < (4LL << 20);
    if (k3_vec_ok<TP, T>(pred, x_mix, x0, a0, D, grad_x, grad_a) && use_tma_pipeline() && prefer_tma) {
        using Op = WmseFwdBwdOp<TP, T>;
        typename Op::Params p{(const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps,
                              w_x, w_a, go_x, go_a, (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a};
        static const int variant = env_int("SISS_K3_VARIANT", 0);     // A/B knob: ring shape
        if (variant == 1) {   // 3 CTAs/SM x 3 stages
            using O1 = WmseFwdBwdOp<TP, T, 3, 3>;
            typename O1::Params q{p.pred, p.x_mix, p.x0, p.a0, p.ts, p.gamma, p.sigma, p.T_steps, p.w_x, p.w_a, p.go_x, p.go_a,
                                  p.grad_x, p.grad_a, p.row_loss_x, p.row_loss_a};
            return launch_pipe<O1>(q, ws, B, D, W, st);
        }
        if (variant == 2) {   // 1 CTA/SM x 10 stages
            using O2 = WmseFwdBwdOp<TP, T, 1, 10>;
            typename O2::Params q{p.pred, p.x_mix, p.x0, p.a0, p.ts, p.gamma, p.sigma, p.T_steps, p.w_x, p.w_a, p.go_x, p.go_a,
                                  p.grad_x, p.grad_a, p.row_loss_x, p.row_loss_a};
            return launch_pipe<O2>(q, ws, B, D, W, st);
        }
        if (variant == 3) {   // 2 CTAs/SM x 4 stages
            using O3 = WmseFwdBwdOp<TP, T, 2, 4>;
            typename O3::Params q{p.pred, p.x_mix, p.x0, p.a0, p.ts, p.gamma, p.sigma, p.T_steps, p.w_x, p.w_a, p.go_x, p.go_a,
                                  p.grad_x, p.grad_a, p.row_loss_x, p.row_loss_a};
            return launch_pipe<O3>(q, ws, B, D, W, st);
        }
        return launch_pipe<Op>(p, ws, B, D, W, st);
    }
    if (k3_vec_ok<TP, T>(pred, x_mix, x0, a0, D, grad_x, grad_a)) {
        RowSched rt = make_row_sched(B, D, W, kK3Occ);
        static const int over = env_int("SISS_LDG_OVERSUB", 1);
        oversubscribe(rt, over, kK3Vpt);
        wmse_fwd_bwd_kernel<TP, T, W, true><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps, w_x, w_a,
            go_x, go_a, (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a, ws, rt);
    } else {
        RowSched rt = make_row_sched(B, D, 1, kK3Occ);
        wmse_fwd_bwd_kernel<TP, T, 1, false><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps, w_x, w_a,
            go_x, go_a, (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a, ws, rt);
    }
    return (int)cudaGetLastError();
}

template <typename TP, typename T>
static int launch_wmse_fwd(const void* pred, const void* x_mix, const void* x0, const void* a0,
                           const int64_t* ts, const float* gamma, const float* sigma, int T_steps,
                           const float* w_x, const float* w_a, float* loss_x, float* loss_a,
                           float* wloss_x, float* wloss_a, long long B, long long D, cudaStream_t st) {
    constexpr int W = VecTraits<T>::N;
    if (k3_vec_ok<TP, T>(pred, x_mix, x0, a0, D, loss_x, loss_a, wloss_x, wloss_a)) {
        RowSched rt = make_row_sched(B, D, W, kK3Occ);
        wmse_fwd_kernel<TP, T, W, true><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps, w_x, w_a,
            loss_x, loss_a, wloss_x, wloss_a, rt);
    } else {
        RowSched rt = make_row_sched(B, D, 1, kK3Occ);
        wmse_fwd_kernel<TP, T, 1, false><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps, w_x, w_a,
            loss_x, loss_a, wloss_x, wloss_a, rt);
    }
    return (int)cudaGetLastError();
}

template <typename TP, typename T>
static int launch_wmse_bwd(const void* pred, const void* x_mix, const void* x0, const void* a0,
                           const int64_t* ts, const float* gamma, const float* sigma, int T_steps,
                           const float* w_x, const float* w_a, GradOut g1, GradOut g2, GradOut g3, GradOut g4,
                           void* grad_pred, long long B, long long D, cudaStream_t st) {
    constexpr int W = VecTraits<T>::N;
    bool vec = k3_vec_ok<TP, T>(pred, x_mix, x0, a0, D, grad_pred, nullptr);
    const GradOut gs[4] = {g1, g2, g3, g4};
    for (const GradOut& g : gs)
        if (g.p && g.stride != 0) vec = vec && aligned16(g.p);
    if (vec) {
        RowSched rt = make_row_sched(B, D, W, kK3Occ);
        wmse_bwd_kernel<TP, T, W, true><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps, w_x, w_a,
            g1, g2, g3, g4, (TP*)grad_pred, rt);
    } else {
        RowSched rt = make_row_sched(B, D, 1, kK3Occ);
        wmse_bwd_kernel<TP, T, 1, false><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred, (const T*)x_mix, (const T*)x0, (const T*)a0, ts, gamma, sigma, T_steps, w_x, w_a,
            g1, g2, g3, g4, (TP*)grad_pred, rt);
    }
    return (int)cudaGetLastError();
}

template <typename TP, typename TT>
static int launch_dual_mse(const void* pred_x, const void* pred_a, const void* tgt_x, const void* tgt_a,
                           float go_x, float go_a, void* grad_x, void* grad_a, float* row_loss_x,
                           float* row_loss_a, void* workspace, long long B, long long D, cudaStream_t st) {
    // unit width: the wider-count of the two dtypes so both sides use whole 128-bit accesses
    constexpr int NP = VecTraits<TP>::N, NT = VecTraits<TT>::N;
    constexpr int W = NP > NT ? NP : NT;
    RowWorkspace ws = carve_row_workspace(workspace, B);
    const bool vec = (D % W == 0) && aligned16(pred_x) && aligned16(pred_a) && aligned16(tgt_x) &&
                     aligned16(tgt_a) && aligned16(grad_x) && aligned16(grad_a);
    if (vec && use_tma_pipeline()) {
        if (tgt_a == tgt_x) {
            using Op = DualMseOp<TP, TT, true>;
            typename Op::Params p{(const TP*)pred_x, (const TP*)pred_a, (const TT*)tgt_x, (const TT*)tgt_a, go_x, go_a,
                                  (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a};
            return launch_pipe<Op>(p, ws, B, D, W, st);
        }
        using Op = DualMseOp<TP, TT, false>;
        typename Op::Params p{(const TP*)pred_x, (const TP*)pred_a, (const TT*)tgt_x, (const TT*)tgt_a, go_x, go_a,
                              (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a};
        return launch_pipe<Op>(p, ws, B, D, W, st);
    }
    if (vec) {
        RowSched rt = make_row_sched(B, D, W, kK3Occ);
        dual_mse_kernel<TP, TT, W, true><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred_x, (const TP*)pred_a, (const TT*)tgt_x, (const TT*)tgt_a, go_x, go_a,
            (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a, ws, rt);
    } else {
        RowSched rt = make_row_sched(B, D, 1, kK3Occ);
        dual_mse_kernel<TP, TT, 1, false><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred_x, (const TP*)pred_a, (const TT*)tgt_x, (const TT*)tgt_a, go_x, go_a,
            (TP*)grad_x, (TP*)grad_a, row_loss_x, row_loss_a, ws, rt);
    }
    return (int)cudaGetLastError();
}

template <typename TP, typename TT>
static int launch_dual_mse_rng(const void* pred_x, const void* pred_a, const void* tgt_x, RngStream rng,
                               const unsigned long long* d_draw, unsigned long long elem_offset, float go_x, float go_a,
                               void* grad_x, void* grad_a, void* tgt_a_out, float* row_loss_x, float* row_loss_a,
                               void* workspace, long long B, long long D, cudaStream_t st) {
    constexpr int NP = VecTraits<TP>::N, NT = VecTraits<TT>::N;
    constexpr int W = NP > NT ? NP : NT;
    RowWorkspace ws = carve_row_workspace(workspace, B);
    const bool vec = (D % W == 0) && (elem_offset % 4 == 0) && aligned16(pred_x) && aligned16(pred_a) && aligned16(tgt_x) &&
                     aligned16(grad_x) && aligned16(grad_a) && aligned16(tgt_a_out);
    if (vec) {
        RowSched rt = make_row_sched(B, D, W, kK3Occ);
        dual_mse_kernel<TP, TT, W, true, true><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred_x, (const TP*)pred_a, (const TT*)tgt_x, nullptr, go_x, go_a, (TP*)grad_x, (TP*)grad_a,
            row_loss_x, row_loss_a, ws, rt, rng, d_draw, elem_offset, (TP*)tgt_a_out);
    } else {
        RowSched rt = make_row_sched(B, D, 1, kK3Occ);
        dual_mse_kernel<TP, TT, 1, false, true><<<rt.grid, kThreads, 0, st>>>(
            (const TP*)pred_x, (const TP*)pred_a, (const TT*)tgt_x, nullptr, go_x, go_a, (TP*)grad_x, (TP*)grad_a,
            row_loss_x, row_loss_a, ws, rt, rng, d_draw, elem_offset, (TP*)tgt_a_out);
    }
    return (int)cudaGetLastError();
}

static int flat_grid(long long n_items) {
    long long blocks = (n_items + kThreads - 1) / kThreads;
    const long long cap = (long long)cached_sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename TP, typename TT, typename TO>
static int launch_sqerr_fwd(const void* pred, const void* tgt, void* loss, void* scaled, float alpha,
                            long long n, cudaStream_t st) {
    sqerr_fwd_kernel<TP, TT, TO><<<flat_grid(n), kThreads, 0, st>>>(
        (const TP*)pred, (const TT*)tgt, (TO*)loss, (TO*)scaled, alpha, n);
    return (int)cudaGetLastError();
}

template <typename TP, typename TT, typename TO>
static int launch_sqerr_bwd(const void* pred, const void* tgt, const void* go_loss, int s1,
                            const void* go_scaled, int s2, float alpha, void* grad_pred, long long n,
                            cudaStream_t st) {
    sqerr_bwd_kernel<TP, TT, TO><<<flat_grid(n), kThreads, 0, st>>>(
        (const TP*)pred, (const TT*)tgt, (const TO*)go_loss, s1, (const TO*)go_scaled, s2, alpha,
        (TP*)grad_pred, n);
    return (int)cudaGetLastError();
}

}  // namespace siss

using namespace siss;

// pred dtype x latent dtype dispatch. Supported: pred fp32 with any latent dtype (accelerate's
// autocast returns fp32 UNet outputs), or pred == latent dtype (pure 16-bit models).
#define SISS_DISPATCH_PRED_IN(pd, id, FN, ...)                                                     \
    do {                                                                                           \
        if ((pd) == SISS_F32 && (id) == SISS_F32)   return FN<float, float>(__VA_ARGS__);          \
        if ((pd) == SISS_F32 && (id) == SISS_BF16)  return FN<float, __nv_bfloat16>(__VA_ARGS__);  \
        if ((pd) == SISS_F32 && (id) == SISS_F16)   return FN<float, __half>(__VA_ARGS__);         \
        if ((pd) == SISS_BF16 && (id) == SISS_BF16) return FN<__nv_bfloat16, __nv_bfloat16>(__VA_ARGS__); \
        if ((pd) == SISS_F16 && (id) == SISS_F16)   return FN<__half, __half>(__VA_ARGS__);        \
        return SISS_EUNSUPPORTED;                                                                  \
    } while (0)

extern "C" {

int siss_wmse_fwd_bwd(const void* pred, int pred_dtype,
                      const void* x_mix, const void* x0, const void* a0, int dtype,
                      const int64_t* timesteps, const float* gamma, const float* sigma, int T_steps,
                      const float* w_x, const float* w_a, float go_x, float go_a,
                      void* grad_x, void* grad_a, float* row_loss_x, float* row_loss_a,
                      void* workspace, int64_t B, int64_t D, siss_stream_t stream) {
    if (!pred || !x_mix || !x0 || !a0 || !timesteps || !gamma || !sigma || !w_x || !w_a || !grad_x || !grad_a ||
        !row_loss_x || !row_loss_a || !workspace || B < 0 || D < 1 || T_steps < 1)
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_PRED_IN(pred_dtype, dtype, launch_wmse_fwd_bwd, pred, x_mix, x0, a0, timesteps, gamma, sigma,
                          T_steps, w_x, w_a, go_x, go_a, grad_x, grad_a, row_loss_x, row_loss_a, workspace, B, D,
                          (cudaStream_t)stream);
}

int siss_wmse_fwd(const void* pred, int pred_dtype,
                  const void* x_mix, const void* x0, const void* a0, int dtype,
                  const int64_t* timesteps, const float* gamma, const float* sigma, int T_steps,
                  const float* w_x, const float* w_a,
                  float* loss_x, float* loss_a, float* wloss_x, float* wloss_a,
                  int64_t B, int64_t D, siss_stream_t stream) {
    if (!pred || !x_mix || !x0 || !a0 || !timesteps || !gamma || !sigma || !w_x || !w_a || B < 0 || D < 1 ||
        T_steps < 1)
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_PRED_IN(pred_dtype, dtype, launch_wmse_fwd, pred, x_mix, x0, a0, timesteps, gamma, sigma, T_steps,
                          w_x, w_a, loss_x, loss_a, wloss_x, wloss_a, B, D, (cudaStream_t)stream);
}

int siss_wmse_bwd(const void* pred, int pred_dtype,
                  const void* x_mix, const void* x0, const void* a0, int dtype,
                  const int64_t* timesteps, const float* gamma, const float* sigma, int T_steps,
                  const float* w_x, const float* w_a,
                  const float* go_loss_x, int go_loss_x_stride,
                  const float* go_loss_a, int go_loss_a_stride,
                  const float* go_wloss_x, int go_wloss_x_stride,
                  const float* go_wloss_a, int go_wloss_a_stride,
                  void* grad_pred, int64_t B, int64_t D, siss_stream_t stream) {
    if (!pred || !x_mix || !x0 || !a0 || !timesteps || !gamma || !sigma || !w_x || !w_a || !grad_pred || B < 0 ||
        D < 1 || T_steps < 1)
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    GradOut g1{go_loss_x, go_loss_x_stride}, g2{go_loss_a, go_loss_a_stride};
    GradOut g3{go_wloss_x, go_wloss_x_stride}, g4{go_wloss_a, go_wloss_a_stride};
    SISS_DISPATCH_PRED_IN(pred_dtype, dtype, launch_wmse_bwd, pred, x_mix, x0, a0, timesteps, gamma, sigma, T_steps,
                          w_x, w_a, g1, g2, g3, g4, grad_pred, B, D, (cudaStream_t)stream);
}

int siss_dual_mse_fwd_bwd(const void* pred_x, const void* pred_a, int pred_dtype,
                          const void* target_x, const void* target_a, int target_dtype,
                          float go_x, float go_a, void* grad_x, void* grad_a,
                          float* row_loss_x, float* row_loss_a,
                          void* workspace, int64_t B, int64_t D, siss_stream_t stream) {
    if (!pred_x || !pred_a || !target_x || !target_a || !grad_x || !grad_a || !row_loss_x || !row_loss_a ||
        !workspace || B < 0 || D < 1)
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_PRED_IN(pred_dtype, target_dtype, launch_dual_mse, pred_x, pred_a, target_x, target_a, go_x, go_a,
                          grad_x, grad_a, row_loss_x, row_loss_a, workspace, B, D, (cudaStream_t)stream);
}

int siss_dual_mse_rng_fwd_bwd(const void* pred_x, const void* pred_a, int pred_dtype, const void* target_x, int target_dtype,
                              uint64_t seed, uint64_t draw, const uint64_t* d_draw, uint64_t elem_offset,
                              float go_x, float go_a, void* grad_x, void* grad_a, void* target_a_out,
                              float* row_loss_x, float* row_loss_a, void* workspace, int64_t B, int64_t D,
                              siss_stream_t stream) {
    if (!pred_x || !pred_a || !target_x || !grad_x || !grad_a || !row_loss_x || !row_loss_a || !workspace || B < 0 ||
        D < 1 || (draw >> 62))
        return SISS_EINVAL;
    if (B == 0) return SISS_OK;
    SISS_DISPATCH_PRED_IN(pred_dtype, target_dtype, launch_dual_mse_rng, pred_x, pred_a, target_x,
                          make_rng_stream(seed, draw, kRngAux), (const unsigned long long*)d_draw, elem_offset, go_x, go_a,
                          grad_x, grad_a, target_a_out, row_loss_x, row_loss_a, workspace, B, D, (cudaStream_t)stream);
}

// promoted output dtype of (pred, target): fp32 unless both are the same 16-bit type
static int promoted(int pd, int td) {
    if (pd == td) return pd;
    if (pd == SISS_F32 || td == SISS_F32) return SISS_F32;
    return -1;  // bf16 x fp16 -> fp32 in torch; not compiled in
}

int siss_sqerr_fwd(const void* pred, int pred_dtype, const void* target, int target_dtype,
                   void* loss, void* scaled, float alpha, int64_t n, siss_stream_t stream) {
    if (!pred || !target || !loss || n < 0) return SISS_EINVAL;
    if (n == 0) return SISS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int od = promoted(pred_dtype, target_dtype);
    if (od == SISS_F32 && pred_dtype == SISS_F32 && target_dtype == SISS_F32 && n % 4 == 0 && aligned16(pred) &&
        aligned16(target) && aligned16(loss) && aligned16(scaled)) {
        sqerr_fwd_f32v4_kernel<<<flat_grid(n / 4), kThreads, 0, st>>>(
            (const float*)pred, (const float*)target, (float*)loss, (float*)scaled, alpha, n / 4);
        return (int)cudaGetLastError();
    }
    if (pred_dtype == SISS_F32 && target_dtype == SISS_F32) return launch_sqerr_fwd<float, float, float>(pred, target, loss, scaled, alpha, n, st);
    if (pred_dtype == SISS_F32 && target_dtype == SISS_BF16) return launch_sqerr_fwd<float, __nv_bfloat16, float>(pred, target, loss, scaled, alpha, n, st);
    if (pred_dtype == SISS_F32 && target_dtype == SISS_F16) return launch_sqerr_fwd<float, __half, float>(pred, target, loss, scaled, alpha, n, st);
    if (pred_dtype == SISS_BF16 && target_dtype == SISS_F32) return launch_sqerr_fwd<__nv_bfloat16, float, float>(pred, target, loss, scaled, alpha, n, st);
    if (pred_dtype == SISS_F16 && target_dtype == SISS_F32) return launch_sqerr_fwd<__half, float, float>(pred, target, loss, scaled, alpha, n, st);
    if (pred_dtype == SISS_BF16 && target_dtype == SISS_BF16) return launch_sqerr_fwd<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(pred, target, loss, scaled, alpha, n, st);
    if (pred_dtype == SISS_F16 && target_dtype == SISS_F16) return launch_sqerr_fwd<__half, __half, __half>(pred, target, loss, scaled, alpha, n, st);
    return SISS_EUNSUPPORTED;
}

int siss_sqerr_bwd(const void* pred, int pred_dtype, const void* target, int target_dtype,
                   const void* go_loss, int go_loss_stride,
                   const void* go_scaled, int go_scaled_stride, float alpha, int go_dtype,
                   void* grad_pred, int64_t n, siss_stream_t stream) {
    if (!pred || !target || !grad_pred || n < 0 || (!go_loss && !go_scaled)) return SISS_EINVAL;
    if (n == 0) return SISS_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (go_dtype != promoted(pred_dtype, target_dtype)) return SISS_EUNSUPPORTED;
    const int s1 = go_loss_stride, s2 = go_scaled_stride;
    if (pred_dtype == SISS_F32 && target_dtype == SISS_F32) return launch_sqerr_bwd<float, float, float>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    if (pred_dtype == SISS_F32 && target_dtype == SISS_BF16) return launch_sqerr_bwd<float, __nv_bfloat16, float>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    if (pred_dtype == SISS_F32 && target_dtype == SISS_F16) return launch_sqerr_bwd<float, __half, float>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    if (pred_dtype == SISS_BF16 && target_dtype == SISS_F32) return launch_sqerr_bwd<__nv_bfloat16, float, float>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    if (pred_dtype == SISS_F16 && target_dtype == SISS_F32) return launch_sqerr_bwd<__half, float, float>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    if (pred_dtype == SISS_BF16 && target_dtype == SISS_BF16) return launch_sqerr_bwd<__nv_bfloat16, __nv_bfloat16, __nv_bfloat16>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    if (pred_dtype == SISS_F16 && target_dtype == SISS_F16) return launch_sqerr_bwd<__half, __half, __half>(pred, target, go_loss, s1, go_scaled, s2, alpha, grad_pred, n, st);
    return SISS_EUNSUPPORTED;
}

}  // extern "C"
