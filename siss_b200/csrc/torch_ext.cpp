// Thin torch extension over the C ABI (include/siss_b200.h): `torch.ops.siss_b200.*`.
//
// BASELINE.json's north_star asks for "hand-written sm_100a CUDA kernels [called] through a thin C-ABI torch
// extension". This file is that layer for the ops on the per-micro-step / per-optimiser-step hot path: each op takes
// tensors, validates them, allocates its outputs, fetches the per-(device, stream) workspace and forwards RAW POINTERS
// to the same `extern "C"` entry points a C / cgo / ctypes caller binds — no kernel code, no arithmetic here. It
// replaces the per-call Python marshalling of siss_b200/ops.py (dtype / device checks, `c_void_p` construction,
// output allocation: 15-35 us per call through ctypes, measured in round 1) with one dispatcher call.
// The ctypes table in siss_b200/_lib.py remains the second, documented binding (and the only one for the cold ops:
// membership metric, multi-tensor K4, peer-memory exchange).
//
// Built by siss_b200/build.py::build_torch_ext with g++ against the torch headers; links libsiss_b200.so by $ORIGIN.

#include <ATen/ATen.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/library.h>

#include <map>
#include <mutex>
#include <tuple>
#include <utility>

#include "../../include/siss_b200.h"

namespace {

using at::Tensor;

int dt_code(const Tensor& t) {
    switch (t.scalar_type()) {
        case at::kFloat: return SISS_F32;
        case at::kBFloat16: return SISS_BF16;
        case at::kHalf: return SISS_F16;
        default: TORCH_CHECK(false, "siss_b200: unsupported dtype ", t.scalar_type(), "; float32 / bfloat16 / float16 only");
    }
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, " failed with code ", rc, ": ", siss_error_string(rc));
}

siss_stream_t cur_stream() { return (siss_stream_t)c10::cuda::getCurrentCUDAStream().stream(); }

void need_cuda_same(std::initializer_list<const Tensor*> ts) {
    const Tensor* first = nullptr;
    for (const Tensor* t : ts) {
        if (!t->defined()) continue;
        TORCH_CHECK(t->is_cuda(), "siss_b200 ops need CUDA tensors on a B200; there is no CPU path");
        if (!first) first = t;
        else TORCH_CHECK(t->device() == first->device(), "tensors on different devices: ", first->device(), " vs ", t->device());
    }
    if (first)
        TORCH_CHECK(first->device().index() == c10::cuda::current_device(), "tensors live on ", first->device(),
                    " but the current CUDA device is cuda:", (int)c10::cuda::current_device());
}

struct Rows { int64_t B, D; };
Rows rows_of(const Tensor& t) {
    const int64_t B = t.dim() ? t.size(0) : 1;
    return {B, B ? t.numel() / B : 0};
}

Tensor timesteps(const Tensor& ts, int64_t B, const at::Device& dev) {
    Tensor t = ts;
    if (t.scalar_type() != at::kLong) t = t.to(at::kLong);
    if (t.dim() == 0) t = t.expand({B});
    if (t.device() != dev) t = t.to(dev, /*non_blocking=*/true);
    return t.contiguous();
}

Tensor table(const Tensor& tab, const at::Device& dev) {
    if (tab.device() != dev || tab.scalar_type() != at::kFloat || !tab.is_contiguous())
        return tab.to(dev, at::kFloat).contiguous();
    return tab;
}

Tensor keep_mask(const Tensor& keep, int64_t B, const at::Device& dev) {
    TORCH_CHECK(keep.dim() == 1 && keep.size(0) == B, "keep_mask must have shape (", B, ",)");
    Tensor k = keep;
    if (k.scalar_type() == at::kBool) k = k.to(at::kByte);
    TORCH_CHECK(k.scalar_type() == at::kByte, "keep_mask must be bool or uint8");
    // the reference draws the mask on the CPU (losses/ddpm_deletion_loss.py:18): B bytes host -> device
    if (k.device() != dev) k = k.to(dev, /*non_blocking=*/true);
    return k.contiguous();
}

// one workspace per (device, stream): zeroed once, the kernels leave it clean
std::mutex g_ws_mutex;
std::map<std::pair<int, void*>, std::pair<int64_t, Tensor>> g_row_ws;
std::map<std::pair<int, void*>, Tensor> g_norm_ws;

void* row_workspace(const at::Device& dev, int64_t B) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    const auto key = std::make_pair((int)dev.index(), (void*)cur_stream());
    auto it = g_row_ws.find(key);
    if (it == g_row_ws.end() || it->second.first < B) {
        const int64_t cap = B > 64 ? B : 64;
        Tensor ws = at::zeros({siss_row_workspace_bytes(cap)}, at::TensorOptions().dtype(at::kByte).device(dev));
        it = g_row_ws.insert_or_assign(key, std::make_pair(cap, ws)).first;
    }
    return it->second.second.data_ptr();
}

void* norm_workspace(const at::Device& dev) {
    std::lock_guard<std::mutex> lock(g_ws_mutex);
    const auto key = std::make_pair((int)dev.index(), (void*)cur_stream());
    auto it = g_norm_ws.find(key);
    if (it == g_norm_ws.end())
        it = g_norm_ws.emplace(key, at::zeros({siss_norm3_workspace_bytes()}, at::TensorOptions().dtype(at::kByte).device(dev))).first;
    return it->second.data_ptr();
}

const void* ptr_or_null(const c10::optional<Tensor>& t) { return (t.has_value() && t->defined()) ? t->data_ptr() : nullptr; }

// ------------------------------------------------------------------------------------------------ K1
Tensor add_noise(const Tensor& x0_, const Tensor& noise_, const Tensor& ts_, const Tensor& ac_) {
    need_cuda_same({&x0_, &noise_});
    TORCH_CHECK(noise_.sizes() == x0_.sizes() && noise_.scalar_type() == x0_.scalar_type(),
                "noise must have the shape and dtype of the samples");
    const Tensor x0 = x0_.contiguous(), noise = noise_.contiguous();
    const Rows r = rows_of(x0);
    const Tensor ts = timesteps(ts_, r.B, x0.device()), ac = table(ac_, x0.device());
    Tensor out = at::empty_like(x0);
    if (out.numel() == 0) return out;
    check_rc(siss_add_noise(x0.data_ptr(), noise.data_ptr(), ts.data_ptr<int64_t>(), ac.data_ptr<float>(), (int)ac.numel(),
                            out.data_ptr(), r.B, r.D, dt_code(x0), cur_stream()), "siss_add_noise");
    return out;
}

std::tuple<Tensor, Tensor> add_noise_pair(const Tensor& x0_, const Tensor& a0_, const Tensor& noise_, const Tensor& ts_,
                                          const Tensor& ac_) {
    need_cuda_same({&x0_, &a0_, &noise_});
    TORCH_CHECK(x0_.sizes() == a0_.sizes() && x0_.sizes() == noise_.sizes() && x0_.scalar_type() == a0_.scalar_type() &&
                x0_.scalar_type() == noise_.scalar_type(), "x0, a0 and noise must share shape and dtype");
    const Tensor x0 = x0_.contiguous(), a0 = a0_.contiguous(), noise = noise_.contiguous();
    const Rows r = rows_of(x0);
    const Tensor ts = timesteps(ts_, r.B, x0.device()), ac = table(ac_, x0.device());
    Tensor xt_x = at::empty_like(x0), xt_a = at::empty_like(a0);
    if (xt_x.numel() == 0) return {xt_x, xt_a};
    check_rc(siss_add_noise_pair(x0.data_ptr(), a0.data_ptr(), noise.data_ptr(), ts.data_ptr<int64_t>(), ac.data_ptr<float>(),
                                 (int)ac.numel(), xt_x.data_ptr(), xt_a.data_ptr(), r.B, r.D, dt_code(x0), cur_stream()),
             "siss_add_noise_pair");
    return {xt_x, xt_a};
}

// ------------------------------------------------------------------------------------------------ K2 / K1oK2
using Five = std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor>;

Five mixture_weights(const Tensor& xt_x_, const Tensor& xt_a_, const Tensor& x0_, const Tensor& a0_, const Tensor& keep_,
                     const Tensor& ts_, const Tensor& gamma_, const Tensor& sigma_, double lambd) {
    need_cuda_same({&xt_x_, &xt_a_, &x0_, &a0_});
    TORCH_CHECK(xt_x_.sizes() == xt_a_.sizes() && xt_x_.sizes() == x0_.sizes() && x0_.sizes() == a0_.sizes(),
                "noisy/original keep/forget batches must share one shape");
    TORCH_CHECK(xt_x_.scalar_type() == xt_a_.scalar_type() && xt_x_.scalar_type() == x0_.scalar_type() &&
                x0_.scalar_type() == a0_.scalar_type(), "noisy/original keep/forget batches must share one dtype");
    const Tensor xt_x = xt_x_.contiguous(), xt_a = xt_a_.contiguous(), x0 = x0_.contiguous(), a0 = a0_.contiguous();
    const Rows r = rows_of(x0);
    const at::Device dev = x0.device();
    const Tensor ts = timesteps(ts_, r.B, dev), keep = keep_mask(keep_, r.B, dev);
    const Tensor g = table(gamma_, dev), s = table(sigma_, dev);
    Tensor x_mix = at::empty_like(xt_x);
    Tensor small = at::empty({4, r.B}, x0.options().dtype(at::kFloat));
    float* sp = small.data_ptr<float>();
    check_rc(siss_mixture_weights(xt_x.data_ptr(), xt_a.data_ptr(), x0.data_ptr(), a0.data_ptr(), keep.data_ptr<uint8_t>(),
                                  ts.data_ptr<int64_t>(), g.data_ptr<float>(), s.data_ptr<float>(), (int)g.numel(), lambd,
                                  x_mix.data_ptr(), sp, sp + r.B, sp + 2 * r.B, sp + 3 * r.B, row_workspace(dev, r.B), r.B,
                                  r.D, dt_code(x0), cur_stream()), "siss_mixture_weights");
    return {x_mix, small.select(0, 0), small.select(0, 1), small.select(0, 2), small.select(0, 3)};
}

Five add_noise_mixture(const Tensor& x0_, const Tensor& a0_, const Tensor& noise_, const Tensor& keep_, const Tensor& ts_,
                       const Tensor& ac_, const Tensor& gamma_, const Tensor& sigma_, double lambd) {
    need_cuda_same({&x0_, &a0_, &noise_});
    TORCH_CHECK(x0_.sizes() == a0_.sizes() && x0_.sizes() == noise_.sizes() && x0_.scalar_type() == a0_.scalar_type() &&
                x0_.scalar_type() == noise_.scalar_type(), "x0, a0 and noise must share shape and dtype");
    const Tensor x0 = x0_.contiguous(), a0 = a0_.contiguous(), noise = noise_.contiguous();
    const Rows r = rows_of(x0);
    const at::Device dev = x0.device();
    const Tensor ts = timesteps(ts_, r.B, dev), keep = keep_mask(keep_, r.B, dev);
    const Tensor ac = table(ac_, dev), g = table(gamma_, dev), s = table(sigma_, dev);
    Tensor x_mix = at::empty_like(x0);
    Tensor small = at::empty({4, r.B}, x0.options().dtype(at::kFloat));
    float* sp = small.data_ptr<float>();
    check_rc(siss_add_noise_mixture(x0.data_ptr(), a0.data_ptr(), noise.data_ptr(), keep.data_ptr<uint8_t>(),
                                    ts.data_ptr<int64_t>(), ac.data_ptr<float>(), g.data_ptr<float>(), s.data_ptr<float>(),
                                    (int)g.numel(), lambd, x_mix.data_ptr(), sp, sp + r.B, sp + 2 * r.B, sp + 3 * r.B,
                                    row_workspace(dev, r.B), r.B, r.D, dt_code(x0), cur_stream()), "siss_add_noise_mixture");
    return {x_mix, small.select(0, 0), small.select(0, 1), small.select(0, 2), small.select(0, 3)};
}

// ------------------------------------------------------------------------------------------------ K3 / dual MSE
using Four = std::tuple<Tensor, Tensor, Tensor, Tensor>;

Four wmse_fwd_bwd(const Tensor& pred_, const Tensor& x_mix_, const Tensor& x0_, const Tensor& a0_, const Tensor& ts_,
                  const Tensor& gamma_, const Tensor& sigma_, const Tensor& w_x_, const Tensor& w_a_, double go_x, double go_a) {
    need_cuda_same({&pred_, &x_mix_, &x0_, &a0_, &w_x_, &w_a_});
    TORCH_CHECK(pred_.sizes() == x_mix_.sizes() && x_mix_.sizes() == x0_.sizes() && x0_.sizes() == a0_.sizes(),
                "pred, x_mix, x0, a0 must share one shape");
    TORCH_CHECK(x_mix_.scalar_type() == x0_.scalar_type() && x0_.scalar_type() == a0_.scalar_type(),
                "x_mix, x0, a0 must share one dtype");
    const Tensor pred = pred_.contiguous(), x_mix = x_mix_.contiguous(), x0 = x0_.contiguous(), a0 = a0_.contiguous();
    const Rows r = rows_of(x0);
    const at::Device dev = x0.device();
    const Tensor ts = timesteps(ts_, r.B, dev), g = table(gamma_, dev), s = table(sigma_, dev);
    const Tensor w_x = w_x_.to(at::kFloat).contiguous(), w_a = w_a_.to(at::kFloat).contiguous();
    Tensor grad_x = at::empty_like(pred), grad_a = at::empty_like(pred);
    Tensor rws = at::empty({2, r.B}, x0.options().dtype(at::kFloat));
    float* rp = rws.data_ptr<float>();
    check_rc(siss_wmse_fwd_bwd(pred.data_ptr(), dt_code(pred), x_mix.data_ptr(), x0.data_ptr(), a0.data_ptr(), dt_code(x0),
                               ts.data_ptr<int64_t>(), g.data_ptr<float>(), s.data_ptr<float>(), (int)g.numel(),
                               w_x.data_ptr<float>(), w_a.data_ptr<float>(), (float)go_x, (float)go_a, grad_x.data_ptr(),
                               grad_a.data_ptr(), rp, rp + r.B, row_workspace(dev, r.B), r.B, r.D, cur_stream()),
             "siss_wmse_fwd_bwd");
    return {grad_x, grad_a, rws.select(0, 0), rws.select(0, 1)};
}

Four dual_mse_fwd_bwd(const Tensor& pred_x_, const Tensor& pred_a_, const Tensor& target_x_, const Tensor& target_a_,
                      double go_x, double go_a) {
    need_cuda_same({&pred_x_, &pred_a_, &target_x_, &target_a_});
    TORCH_CHECK(pred_x_.sizes() == pred_a_.sizes() && pred_a_.sizes() == target_x_.sizes() &&
                target_x_.sizes() == target_a_.sizes(), "preds and targets must share one shape");
    TORCH_CHECK(pred_x_.scalar_type() == pred_a_.scalar_type() && target_x_.scalar_type() == target_a_.scalar_type(),
                "pred_x/pred_a and target_x/target_a must pairwise share dtypes");
    const bool same = target_a_.is_same(target_x_) || target_a_.data_ptr() == target_x_.data_ptr();
    const Tensor pred_x = pred_x_.contiguous(), pred_a = pred_a_.contiguous(), target_x = target_x_.contiguous();
    const Tensor target_a = same ? target_x : target_a_.contiguous();
    const Rows r = rows_of(pred_x);
    const at::Device dev = pred_x.device();
    Tensor grad_x = at::empty_like(pred_x), grad_a = at::empty_like(pred_a);
    Tensor rws = at::empty({2, r.B}, pred_x.options().dtype(at::kFloat));
    float* rp = rws.data_ptr<float>();
    check_rc(siss_dual_mse_fwd_bwd(pred_x.data_ptr(), pred_a.data_ptr(), dt_code(pred_x), target_x.data_ptr(),
                                   target_a.data_ptr(), dt_code(target_x), (float)go_x, (float)go_a, grad_x.data_ptr(),
                                   grad_a.data_ptr(), rp, rp + r.B, row_workspace(dev, r.B), r.B, r.D, cur_stream()),
             "siss_dual_mse_fwd_bwd");
    return {grad_x, grad_a, rws.select(0, 0), rws.select(0, 1)};
}

// ------------------------------------------------------------------------------------------------ K4
void norm3_(const Tensor& g_x, const Tensor& g_a, Tensor out) {
    need_cuda_same({&g_x, &g_a, &out});
    TORCH_CHECK(g_x.scalar_type() == at::kFloat && g_a.scalar_type() == at::kFloat, "gradient buffers must be float32");
    TORCH_CHECK(g_x.numel() == g_a.numel() && g_x.is_contiguous() && g_a.is_contiguous(),
                "g_x and g_a must be contiguous and equally sized");
    TORCH_CHECK(out.scalar_type() == at::kDouble && out.numel() >= 3 && out.is_contiguous(), "out must be float64[3]");
    check_rc(siss_norm3(g_x.data_ptr<float>(), g_a.data_ptr<float>(), g_x.numel(), out.data_ptr<double>(),
                        norm_workspace(g_x.device()), cur_stream()), "siss_norm3");
}

void combine_(const Tensor& g_x, const Tensor& g_a, const Tensor& sums3, int64_t mode, double value, double max_norm,
              bool inf_guard, Tensor out, Tensor stats) {
    need_cuda_same({&g_x, &g_a, &sums3, &out, &stats});
    TORCH_CHECK(g_x.scalar_type() == at::kFloat && g_a.scalar_type() == at::kFloat && sums3.scalar_type() == at::kDouble,
                "gradient buffers must be float32 and sums3 float64");
    TORCH_CHECK(out.scalar_type() == at::kFloat && out.numel() == g_x.numel() && g_a.numel() == g_x.numel() &&
                g_x.is_contiguous() && g_a.is_contiguous() && out.is_contiguous(), "out must match the gradient buffers");
    TORCH_CHECK(stats.scalar_type() == at::kFloat && stats.numel() >= 5 && stats.is_contiguous(), "stats must be float32[5]");
    check_rc(siss_combine(g_x.data_ptr<float>(), g_a.data_ptr<float>(), out.data_ptr<float>(), g_x.numel(),
                          sums3.data_ptr<double>(), (int)mode, (float)value, (float)max_norm, inf_guard ? 1 : 0,
                          stats.data_ptr<float>(), cur_stream()), "siss_combine");
}

// ------------------------------------------------------------------------------------------------ statistics
void batch_stats_(const c10::optional<Tensor>& rl_x, const c10::optional<Tensor>& rl_a, const c10::optional<Tensor>& w_x,
                  const c10::optional<Tensor>& w_a, int64_t elems_per_sample, Tensor out) {
    int64_t B = -1;
    for (const auto* t : {&rl_x, &rl_a, &w_x, &w_a}) {
        if (!t->has_value() || !(*t)->defined()) continue;
        const Tensor& v = **t;
        TORCH_CHECK(v.is_cuda(), "siss_b200 ops need CUDA tensors on a B200; there is no CPU path");
        if (B < 0) B = v.numel();
        TORCH_CHECK(v.scalar_type() == at::kFloat && v.numel() == B && v.is_contiguous(),
                    "batch_stats inputs must be contiguous float32 vectors of one length");
    }
    TORCH_CHECK(B >= 0, "batch_stats needs at least one input");
    TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kFloat && out.numel() >= 16 && out.is_contiguous(),
                "out must be float32[16] on the device");
    check_rc(siss_batch_stats((const float*)ptr_or_null(rl_x), (const float*)ptr_or_null(rl_a), (const float*)ptr_or_null(w_x),
                              (const float*)ptr_or_null(w_a), B, elems_per_sample, out.data_ptr<float>(), cur_stream()),
             "siss_batch_stats");
}

}  // namespace

TORCH_LIBRARY(siss_b200, m) {
    m.def("add_noise(Tensor x0, Tensor noise, Tensor timesteps, Tensor alphas_cumprod) -> Tensor");
    m.def("add_noise_pair(Tensor x0, Tensor a0, Tensor noise, Tensor timesteps, Tensor alphas_cumprod) -> (Tensor, Tensor)");
    m.def("mixture_weights(Tensor xt_x, Tensor xt_a, Tensor x0, Tensor a0, Tensor keep_mask, Tensor timesteps, Tensor gamma, "
          "Tensor sigma, float lambd) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
    m.def("add_noise_mixture(Tensor x0, Tensor a0, Tensor noise, Tensor keep_mask, Tensor timesteps, Tensor alphas_cumprod, "
          "Tensor gamma, Tensor sigma, float lambd) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
    m.def("wmse_fwd_bwd(Tensor pred, Tensor x_mix, Tensor x0, Tensor a0, Tensor timesteps, Tensor gamma, Tensor sigma, "
          "Tensor w_x, Tensor w_a, float go_x, float go_a) -> (Tensor, Tensor, Tensor, Tensor)");
    m.def("dual_mse_fwd_bwd(Tensor pred_x, Tensor pred_a, Tensor target_x, Tensor target_a, float go_x, float go_a) -> "
          "(Tensor, Tensor, Tensor, Tensor)");
    m.def("norm3_(Tensor g_x, Tensor g_a, Tensor(a!) out) -> ()");
    m.def("combine_(Tensor g_x, Tensor g_a, Tensor sums3, int mode, float value, float max_norm, bool inf_guard, "
          "Tensor(a!) out, Tensor(b!) stats) -> ()");
    m.def("batch_stats_(Tensor? row_loss_x, Tensor? row_loss_a, Tensor? w_x, Tensor? w_a, int elems_per_sample, "
          "Tensor(a!) out) -> ()");
}

TORCH_LIBRARY_IMPL(siss_b200, CUDA, m) {
    m.impl("add_noise", &add_noise);
    m.impl("add_noise_pair", &add_noise_pair);
    m.impl("mixture_weights", &mixture_weights);
    m.impl("add_noise_mixture", &add_noise_mixture);
    m.impl("wmse_fwd_bwd", &wmse_fwd_bwd);
    m.impl("dual_mse_fwd_bwd", &dual_mse_fwd_bwd);
    m.impl("norm3_", &norm3_);
    m.impl("combine_", &combine_);
}

// batch_stats_ takes only optional inputs: register it for every backend key it can be dispatched on
TORCH_LIBRARY_IMPL(siss_b200, CompositeExplicitAutograd, m) {
    m.impl("batch_stats_", &batch_stats_);
}
