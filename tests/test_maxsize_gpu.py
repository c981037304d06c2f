"""GPU: maximum sizes — more than 2**31 elements (byte offsets beyond 2**32) on the vector / TMA path and more than
2**31 UNITS on the scalar path (odd D), for the row kernels and the flat gradient kernels. The oracle cannot run at
this size; the property used is row independence: any block of rows of the big launch must be BIT-IDENTICAL to a
small launch on just those rows (first rows, rows straddling the 2**31 / 2**32 boundaries, last rows), plus the
float64 identities of test_fullsize_gpu.py on those blocks. Needs ~45 GB of the B200's 180 GB; skipped on smaller GPUs."""
import pytest
import torch

pytestmark = pytest.mark.gpu

D_VEC, D_ODD = 3 * 256 * 256, 3 * 256 * 256 + 1


@pytest.fixture(scope="module")
def dev(cuda_device):
    from siss_b200 import _lib
    _lib.load()
    if torch.cuda.get_device_properties(cuda_device).total_memory < 100 * (1 << 30):
        pytest.skip("needs a >=100 GB GPU")
    return cuda_device


def _blocks(B, D):
    """Row blocks to re-run in isolation: start, around element 2**31, (around byte 2**32 is the same rows for 2-byte
    types), end."""
    r31 = (1 << 31) // D
    return [(0, 3), (r31 - 1, r31 + 2), (B - 3, B)]


@pytest.mark.parametrize("dtype,D", [(torch.bfloat16, D_VEC), (torch.float32, D_ODD)])
def test_row_kernels_beyond_2_31(dtype, D, dev):
    from siss_b200 import ops
    from siss_b200.scheduler import SissDDPMScheduler
    B = (1 << 31) // D + 300                        # > 2**31 elements; with odd D also > 2**31 scalar units
    assert B * D > (1 << 31)
    sched = SissDDPMScheduler(); ac = sched.alphas_cumprod.to(dev); gamma, sigma = sched.gamma_sigma(dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x0 = torch.empty(B, D, dtype=dtype, device=dev).uniform_(-1, 1, generator=g)
    a0 = torch.empty(B, D, dtype=dtype, device=dev).uniform_(-1, 1, generator=g)
    nz = torch.empty(B, D, dtype=dtype, device=dev).normal_(generator=g)
    t = torch.randint(0, 1000, (B,), device=dev, generator=g)
    keep = torch.rand(B, device=dev, generator=g) > 0.5
    x_mix, d_x, d_a, w_x, w_a = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
    xt_x, xt_a = ops.add_noise_pair(x0, a0, nz, t, ac)
    for lo, hi in _blocks(B, D):
        s = slice(lo, hi)
        part = ops.add_noise_mixture(x0[s], a0[s], nz[s], keep[s], t[s], ac, gamma, sigma, 0.5)
        assert torch.equal(part[0], x_mix[s]), (lo, hi)
        # row sums are reduced in a launch-dependent order -> fp32 summation tolerance, weights follow
        torch.testing.assert_close(part[1], d_x[s], rtol=1e-5, atol=0)
        torch.testing.assert_close(part[2], d_a[s], rtol=1e-5, atol=0)
        px, pa = ops.add_noise_pair(x0[s], a0[s], nz[s], t[s], ac)
        assert torch.equal(px, xt_x[s]) and torch.equal(pa, xt_a[s])
        sel = torch.where(keep[s, None], px, pa)
        assert torch.equal(part[0], sel)                                   # x_mix rows are the selected noisy rows
        # float64 re-evaluation of d_x on the device for these rows
        gam, sig = gamma[t[s]].double()[:, None], sigma[t[s]].double()[:, None]
        want = ((x_mix[s].double() - gam * x0[s].double()) ** 2).sum(1) / (2 * sig[:, 0] ** 2)
        torch.testing.assert_close(d_x[s].double(), want, rtol=2e-5, atol=0)
    del xt_x, xt_a
    # K3 on the same tensors
    pred = torch.empty(B, D, dtype=torch.float32, device=dev).normal_(generator=g)
    g_x, g_a, rl_x, rl_a = ops.wmse_fwd_bwd(pred, x_mix, x0, a0, t, gamma, sigma, w_x, w_a, 1 / 64, 1 / 64)
    for lo, hi in _blocks(B, D):
        s = slice(lo, hi)
        p = ops.wmse_fwd_bwd(pred[s], x_mix[s], x0[s], a0[s], t[s], gamma, sigma, w_x[s], w_a[s], 1 / 64, 1 / 64)
        assert torch.equal(p[0], g_x[s]) and torch.equal(p[1], g_a[s]), (lo, hi)
        torch.testing.assert_close(p[2], rl_x[s], rtol=1e-5, atol=0)
        torch.testing.assert_close(p[3], rl_a[s], rtol=1e-5, atol=0)
    assert torch.isfinite(rl_x).all() and torch.isfinite(g_a[-1]).all()


def test_flat_gradient_kernels_beyond_2_31(dev):
    """K4a / K4b / fused AdamW over 2**31 + 1027 parameters (SD-1.4 has 0.86e9; this is the indexing limit test)."""
    from siss_b200 import _lib, ops
    n = (1 << 31) + 1027
    g = torch.Generator(device=dev).manual_seed(2)
    gx = torch.empty(n, device=dev).normal_(generator=g).mul_(1e-3)
    ga = torch.empty(n, device=dev).normal_(generator=g).mul_(1e-3)
    sums = ops.norm3(gx, ga)
    want = torch.stack([torch.linalg.vector_norm(gx, dtype=torch.float64) ** 2,
                        torch.linalg.vector_norm(ga, dtype=torch.float64) ** 2])
    torch.testing.assert_close(sums[:2], want, rtol=1e-9, atol=0)
    out, stats = ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0)
    n_a = sums[1].sqrt().item()
    s, clip = stats[2].item(), stats[4].item()
    assert abs(s - 500.0 / n_a) < 1e-5 * s
    for lo, hi in [(0, 4096), ((1 << 31) - 2048, (1 << 31) + 1027)]:       # both ends incl. the ragged tail
        ref = (gx[lo:hi] - stats[2] * ga[lo:hi]) * stats[4]                 # eager fp32: mul, sub, mul — the kernel's order
        assert torch.equal(out[lo:hi], ref), (lo, hi)
    torch.testing.assert_close(torch.linalg.vector_norm(out, dtype=torch.float64).item(), 1.0, rtol=1e-5, atol=0)
