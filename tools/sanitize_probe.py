#!/usr/bin/env python
"""Workload for compute-sanitizer: every kernel family once, at shapes that exercise multi-stage TMA rings
(several ring wrap-arounds per CTA), cross-CTA row tickets, the scalar fallback path and the flat /
multi-tensor / fused-optimiser K4 variants. Kept small: the sanitizer slows kernels 10-100x."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from siss_b200 import _lib, ops
from siss_b200.rng import DeviceRng
from siss_b200.scheduler import SissDDPMScheduler

dev = torch.device("cuda", 0)
sched = SissDDPMScheduler(); ac = sched.alphas_cumprod.to(dev); gamma, sigma = sched.gamma_sigma(dev)
for dt in (torch.bfloat16, torch.float32):
    for shape in [(6, 3, 256, 256), (40, 1, 28, 28), (3, 3, 7, 5)]:
        B = shape[0]
        x0 = (torch.rand(shape, device=dev) * 2 - 1).to(dt); a0 = (torch.rand(shape, device=dev) * 2 - 1).to(dt)
        nz = torch.randn(shape, device=dev).to(dt); pred = torch.randn(shape, device=dev)
        t = torch.randint(0, 1000, (B,), device=dev); keep = torch.rand(B, device=dev) > 0.5
        xt_x, xt_a = ops.add_noise_pair(x0, a0, nz, t, ac)
        ops.add_noise(x0, nz, t, ac)
        m = ops.mixture_weights(xt_x, xt_a, x0, a0, keep, t, gamma, sigma, 0.5)
        f = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
        ops.wmse_fwd_bwd(pred, f[0], x0, a0, t, gamma, sigma, f[3], f[4], 1 / 64, 1 / 64)
        lx = ops.wmse_fwd(pred, f[0], x0, a0, t, gamma, sigma, f[3], f[4])
        ops.wmse_bwd(pred, f[0], x0, a0, t, gamma, sigma, f[3], f[4], go_wloss_x=torch.full((), 0.1, device=dev).expand(shape))
        ops.dual_mse_fwd_bwd(pred, pred, nz, nz, 0.1, 0.1)
        l, s = ops.sqerr_fwd(pred, nz, alpha=-1.5)
        ops.sqerr_bwd(pred, nz, go_loss=torch.ones_like(l))
        ops.batch_stats(m[1], m[2], m[3], m[4], shape[1] * shape[2] * shape[3])
        # opt-in device RNG: randn, per-row draws, eps generated inside K1oK2 (TMA ring with 2 streams / LDG / scalar path)
        rng = DeviceRng(seed=3, row_offset=1)
        rng.randn(shape, dt, dev, draw=2)
        rng.draw_rows(B, dev, t_range=(0, 1000), lambd=0.5)
        ops.add_noise_mixture_rng(x0, a0, keep, t, ac, gamma, sigma, 0.5, 3, 2, elem_offset=x0[0].numel(), want_noise=True)
        ops.add_noise_mixture_rng(x0, a0, keep, t, ac, gamma, sigma, 0.5, 3, 2, elem_offset=3)          # unaligned -> scalar path
        # membership metric: expanded-row slices of an I x N_n grid, rows split over CTAs
        mx, ma = ops.membership_add_noise(x0, a0, nz[:2], 437, ac, 1, 2 * B - 1)
        ops.membership_sqerr(mx.float(), ma.float(), nz[:2], 1)
n = (1 << 20) + 3
gx, ga = torch.randn(n, device=dev), torch.randn(n, device=dev)
sums = ops.norm3(gx, ga)
ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 5.0, 1.0)
ops.combine(gx, ga, sums, _lib.SISS_COMBINE_ERASEDIFF, 0.05, 1.0, out=gx)
xs = [torch.randn(k, device=dev) for k in (5, 4096, 9000)]; as_ = [torch.randn(k, device=dev) for k in (5, 4096, 9000)]
ops.MultiTensorPlan(xs, as_).combine(_lib.SISS_COMBINE_SCALING_NORM, 5.0, 1.0)
from siss_b200.grad_combine import GradCombiner
from siss_b200.optim import FusedCombineAdamW
net = torch.nn.Linear(300, 70).to(dev)
comb = GradCombiner(net.parameters()); opt = FusedCombineAdamW(comb, lr=1e-3)
comb.begin_x(); net(torch.randn(8, 300, device=dev)).square().mean().backward()
comb.begin_a(); net(torch.randn(8, 300, device=dev)).square().mean().backward()
opt.step(scaling_norm=5.0)
opt2 = FusedCombineAdamW(comb, lr=1e-3, ema=dict(decay=0.99), device_schedule=True)       # EMA + device-side {lr, decay}
comb.begin_x(); net(torch.randn(8, 300, device=dev)).square().mean().backward()
opt2.step(single_term=True); opt2.set_schedule(lr=5e-4)
comb.begin_x(); net(torch.randn(8, 300, device=dev)).square().mean().backward()
opt2.step(single_term=True)
torch.cuda.synchronize()
print("sanitize probe done")
