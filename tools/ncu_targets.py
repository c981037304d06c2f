#!/usr/bin/env python
"""One warm + one profiled launch of every kernel that round 1 left without an ncu capture, at the delete_celeb shape
(B = 64 x 3x256x256 bf16 latents, fp32 eps_hat, P = 113.67 M): run under
    ncu --set full --clock-control none -k regex:<see tools/ncu_r2.sh> ...
Prints the algorithmic bytes of each launch (JSON) so that the summary can put them beside dram__bytes."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch  # noqa: E402

from siss_b200 import _lib, ops  # noqa: E402
from siss_b200.grad_combine import GradCombiner  # noqa: E402
from siss_b200.metrics.class_membership import MembershipLoss  # noqa: E402,F401
from siss_b200.optim import FusedCombineAdamW  # noqa: E402
from siss_b200.rng import DeviceRng  # noqa: E402
from siss_b200.scheduler import SissDDPMScheduler  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B, shape, D = 64, (64, 3, 256, 256), 3 * 256 * 256
dt, s_in = torch.bfloat16, 2
P = 113_673_220
g = torch.Generator(device=dev).manual_seed(0)
sched = SissDDPMScheduler()
ac = sched.alphas_cumprod.to(dev)
gamma, sigma = sched.gamma_sigma(dev)
x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dt)
a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dt)
nz = torch.randn(shape, device=dev, generator=g).to(dt)
pred, pred2 = torch.randn(shape, device=dev, generator=g), torch.randn(shape, device=dev, generator=g)
t = torch.full((B,), 999, device=dev, dtype=torch.long)
keep = (torch.rand(B, device=dev, generator=g) > 0.5).to(torch.uint8)
N = B * D
alg = {}


def run(name, bytes_per_launch, fn):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    alg[name] = int(bytes_per_launch)


run("siss_add_noise_pair", 5 * s_in * N, lambda: ops.add_noise_pair(x0, a0, nz, t, ac))
run("siss_add_noise", 3 * s_in * N, lambda: ops.add_noise(x0, nz, t, ac))
xt_x, xt_a = ops.add_noise_pair(x0, a0, nz, t, ac)
run("siss_mixture_weights", 4 * s_in * N, lambda: ops.mixture_weights(xt_x, xt_a, x0, a0, keep, t, gamma, sigma, 0.5))
run("siss_add_noise_mixture_rng", 3 * s_in * N, lambda: ops.add_noise_mixture_rng(x0, a0, keep, t, ac, gamma, sigma, 0.5, 42, 7))
run("siss_dual_mse_fwd_bwd", (16 + s_in) * N, lambda: ops.dual_mse_fwd_bwd(pred, pred2, nz, nz, 1 / 64, 1 / 64))
run("siss_dual_mse_rng_fwd_bwd", (16 + s_in) * N, lambda: ops.dual_mse_rng_fwd_bwd(pred, pred2, nz, 1 / 64, 1 / 64, 42, 7))
x_mix, d_x, d_a, w_x, w_a = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
run("siss_wmse_fwd(api-compat)", (20 + 3 * s_in) * N, lambda: ops.wmse_fwd(pred, x_mix, x0, a0, t, gamma, sigma, w_x, w_a))
gx, ga, rl_x, rl_a = ops.wmse_fwd_bwd(pred, x_mix, x0, a0, t, gamma, sigma, w_x, w_a, 1 / 64, 1 / 64)
run("siss_batch_stats", 4 * 4 * B, lambda: ops.batch_stats(rl_x, rl_a, w_x, w_a, D))
run("siss_randn", s_in * N, lambda: DeviceRng(42).randn(shape, dt, dev, draw=3))
run("siss_draw_rows", 9 * B, lambda: DeviceRng(42).draw_rows(B, dev, t_range=(0, 1000), lambd=0.5))
del pred2, xt_x, xt_a, gx, ga
torch.cuda.empty_cache()

# fused combine + AdamW (+EMA): 40 (+8) B/param in one launch
holder = torch.nn.Parameter(torch.randn(P, device=dev) * 1e-2)
comb = GradCombiner([holder])
opt = FusedCombineAdamW(comb, lr=5e-6, betas=(0.95, 0.999), eps=1e-8, weight_decay=1e-6)


def adamw():
    comb.g_x.normal_(0, 1e-3); comb.g_a.normal_(0, 1e-3)
    opt.step(scaling_norm=500.0, max_norm=1.0)


run("siss_combine_adamw", 40 * comb.total, adamw)
del opt, comb, holder
torch.cuda.empty_cache()

# multi-tensor K4 over a ragged list of per-parameter tensors (same total size)
sizes = [P // 64] * 63 + [P - 63 * (P // 64)]
xs = [torch.randn(n, device=dev) * 1e-3 for n in sizes]
as_ = [torch.randn(n, device=dev) * 1e-3 for n in sizes]
plan = ops.MultiTensorPlan(xs, as_)
run("siss_mt_norm3+siss_mt_combine", 20 * P, lambda: plan.combine(_lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0))
del plan, xs, as_
torch.cuda.empty_cache()

# membership-loss metric kernels (metrics/class_membership.py:76-116): 8 images x 8 noises at the celeb shape
I, Nn = 8, 8
mi = (torch.rand((I, 3, 256, 256), device=dev, generator=g) * 2 - 1).to(dt)
mn = torch.randn((Nn, 3, 256, 256), device=dev, generator=g).to(dt)
rows = I * Nn
run("siss_membership_add_noise", (2 * I + Nn) * D * s_in + 2 * rows * D * s_in, lambda: ops.membership_add_noise(mi, mi, mn, 500, ac, 0, rows))
px, pa = torch.randn((rows, 3, 256, 256), device=dev, generator=g), torch.randn((rows, 3, 256, 256), device=dev, generator=g)
run("siss_membership_sqerr", (8 + s_in) * rows * D, lambda: ops.membership_sqerr(px, pa, mn, 0))
print("ALG_BYTES " + json.dumps(alg), flush=True)
