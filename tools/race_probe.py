import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from siss_b200 import ops
from siss_b200.scheduler import SissDDPMScheduler
dev = torch.device("cuda", 0)
B, C, H, W = 64, 3, 256, 256
g = torch.Generator(device=dev).manual_seed(42)
sched = SissDDPMScheduler(); ac = sched.alphas_cumprod.to(dev); gamma, sigma = sched.gamma_sigma(dev)
shape = (B, C, H, W)
x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
nz = torch.randn(shape, device=dev, generator=g).bfloat16()
pred = torch.randn(shape, device=dev, generator=g)
t = torch.full((B,), 999, device=dev, dtype=torch.long)
keep = torch.rand(B, device=dev, generator=g) > 0.5
f = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
f2 = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
print("K1K2 deterministic:", [torch.equal(a, b) for a, b in zip(f, f2)])
x_mix, dx, da, wx, wa = f
args = (x_mix, x0, a0, t, gamma, sigma, wx, wa)
r1 = ops.wmse_fwd_bwd(pred, *args, 1 / 64, 1 / 64)
r2 = ops.wmse_fwd_bwd(pred, *args, 1 / 64, 1 / 64)
print("K3 deterministic:", [torch.equal(a, b) for a, b in zip(r1, r2)])
gs, ss = gamma[t].view(-1, 1, 1, 1), sigma[t].view(-1, 1, 1, 1)
eps_x = (x_mix.float() - gs * x0.float()) / ss
ref_gx = (torch.tensor(1 / 64, device=dev) * wx).view(-1, 1, 1, 1) * (2 * (pred - eps_x))
d = (r1[0] != ref_gx)
print("K3 vs eager mismatches:", d.sum().item(), "of", d.numel())
if d.any():
    idx = d.nonzero()[:5]; print(idx.tolist()); print(r1[0][d][:5], ref_gx[d][:5])
r3 = ops.wmse_fwd_bwd(pred, *args, 2 / 64, 4 / 64)
d2 = (r3[0] != 2 * r1[0])
print("linearity mismatches:", d2.sum().item())
if d2.any():
    print(r3[0][d2][:5], (2 * r1[0])[d2][:5], wx[d2.nonzero()[:5, 0]])
