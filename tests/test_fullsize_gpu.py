"""GPU: BASELINE.json's full sizes (delete_celeb shape, B = 64, bf16 latents; P = 113.67 M gradients).
The CPU oracle would take seconds per case here, so parity is checked through size-independent
properties and an independent float64 re-evaluation ON THE DEVICE with plain torch ops:

  * K1oK2 == K1 followed by K2 (bit-exact), x_mix rows are exactly the selected noisy rows;
  * the weight identity (1-l) w_x + l w_a == 1; d_x, d_a vs float64 torch on device;
  * K3 gradients are linear in the upstream scale and vanish for pred == eps_x / eps_a respectively;
  * row sums vs float64 torch; K4: ||.|| vs torch.linalg.vector_norm(float64), result norm == min(1, .),
    linearity of the combine in (g_x, g_a) at fixed scalars, idempotence of the clip.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

B, C, H, W = 64, 3, 256, 256
P = 113_673_219


@pytest.fixture(scope="module")
def env(cuda_device):
    from siss_b200 import _lib
    from siss_b200.scheduler import SissDDPMScheduler
    _lib.load()
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(42)
    sched = SissDDPMScheduler()
    d = dict(dev=dev, sched=sched, ac=sched.alphas_cumprod.to(dev))
    d["gamma"], d["sigma"] = sched.gamma_sigma(dev)
    shape = (B, C, H, W)
    d["x0"] = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
    d["a0"] = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
    d["noise"] = torch.randn(shape, device=dev, generator=g).bfloat16()
    d["pred"] = torch.randn(shape, device=dev, generator=g)
    d["t"] = torch.full((B,), 999, device=dev, dtype=torch.long)          # delete_celeb.py:593
    d["keep"] = torch.rand(B, device=dev, generator=g) > 0.5
    return d


def test_fused_equals_unfused_and_select(env):
    from siss_b200 import ops
    e = env
    xt_x, xt_a = ops.add_noise_pair(e["x0"], e["a0"], e["noise"], e["t"], e["ac"])
    assert torch.equal(xt_x, ops.add_noise(e["x0"], e["noise"], e["t"], e["ac"]))
    f = ops.add_noise_mixture(e["x0"], e["a0"], e["noise"], e["keep"], e["t"], e["ac"], e["gamma"], e["sigma"], 0.5)
    u = ops.mixture_weights(xt_x, xt_a, e["x0"], e["a0"], e["keep"], e["t"], e["gamma"], e["sigma"], 0.5)
    for a, b in zip(f, u):
        assert torch.equal(a, b)
    k = e["keep"]
    assert torch.equal(f[0][k], xt_x[k]) and torch.equal(f[0][~k], xt_a[~k])
    # eager torch on the device computes the same bf16 x_t (independent of our kernels)
    a = e["ac"].to(torch.bfloat16)[e["t"]]
    ref = (a ** 0.5).view(-1, 1, 1, 1) * e["x0"] + ((1 - a) ** 0.5).view(-1, 1, 1, 1) * e["noise"]
    assert torch.equal(xt_x, ref)
    e["x_mix"], e["d_x"], e["d_a"], e["w_x"], e["w_a"] = f


def test_weights_identity_and_float64_exponents(env):
    e = env
    ident = 0.5 * e["w_x"].double() + 0.5 * e["w_a"].double()
    torch.testing.assert_close(ident, torch.ones_like(ident), rtol=1e-6, atol=1e-6)
    g, s = e["gamma"].double()[e["t"]].view(-1, 1, 1, 1), e["sigma"].double()[e["t"]]
    xm = e["x_mix"].double()
    # same fp32 residuals as the reference (fp32 gamma * x0, fp32 subtraction), float64 accumulation
    rx = (e["x_mix"].float() - e["gamma"][e["t"]].view(-1, 1, 1, 1) * e["x0"].float()).double()
    ra = (e["x_mix"].float() - e["gamma"][e["t"]].view(-1, 1, 1, 1) * e["a0"].float()).double()
    dx, da = (rx ** 2).sum(dim=[1, 2, 3]) / (2 * s ** 2), (ra ** 2).sum(dim=[1, 2, 3]) / (2 * s ** 2)
    torch.testing.assert_close(e["d_x"].double(), dx, rtol=2e-6, atol=0)
    torch.testing.assert_close(e["d_a"].double(), da, rtol=2e-6, atol=0)
    w_x64 = 1 / (0.5 + 0.5 * torch.exp(dx - da))
    torch.testing.assert_close(e["w_x"].double(), w_x64, rtol=1e-4, atol=0)   # directly accumulated difference
    del xm, g


def test_k3_linearity_zero_residual_and_row_sums(env):
    from siss_b200 import ops
    e = env
    args = (e["x_mix"], e["x0"], e["a0"], e["t"], e["gamma"], e["sigma"], e["w_x"], e["w_a"])
    gx1, ga1, rlx, rla = ops.wmse_fwd_bwd(e["pred"], *args, 1 / 64, 1 / 64)
    gx2, ga2, _, _ = ops.wmse_fwd_bwd(e["pred"], *args, 2 / 64, 4 / 64)
    assert torch.equal(gx2, 2 * gx1) and torch.equal(ga2, 4 * ga1)          # powers of two: exactly linear
    g, s = e["gamma"][e["t"]].view(-1, 1, 1, 1), e["sigma"][e["t"]].view(-1, 1, 1, 1)
    eps_x = (e["x_mix"].float() - g * e["x0"].float()) / s
    eps_a = (e["x_mix"].float() - g * e["a0"].float()) / s
    gx0, _, r0, _ = ops.wmse_fwd_bwd(eps_x, *args, 1 / 64, 1 / 64)
    assert gx0.abs().max().item() == 0.0 and r0.abs().max().item() == 0.0      # pred == eps_x: zero loss_x and grad_x
    _, ga0, _, r1 = ops.wmse_fwd_bwd(eps_a, *args, 1 / 64, 1 / 64)
    assert ga0.abs().max().item() == 0.0 and r1.abs().max().item() == 0.0
    torch.testing.assert_close(rlx.double(), ((e["pred"] - eps_x).double() ** 2).sum(dim=[1, 2, 3]), rtol=2e-6, atol=0)
    torch.testing.assert_close(rla.double(), ((e["pred"] - eps_a).double() ** 2).sum(dim=[1, 2, 3]), rtol=2e-6, atol=0)
    ref_gx = (torch.tensor(1 / 64, device=e["dev"]) * e["w_x"]).view(-1, 1, 1, 1) * (2 * (e["pred"] - eps_x))
    assert torch.equal(gx1, ref_gx)                                          # eager on device, same op order


def test_k4_full_size_properties(env):
    from siss_b200 import ops, _lib
    dev = env["dev"]
    g = torch.Generator(device=dev).manual_seed(7)
    gx = torch.randn(P, device=dev, generator=g) * 1e-3
    ga = torch.randn(P, device=dev, generator=g) * 1e-3
    sums = ops.norm3(gx, ga)
    chunks = lambda v, w: sum((a.double() * b.double()).sum() for a, b in zip(v.split(1 << 24), w.split(1 << 24)))
    ref = torch.stack([chunks(gx, gx), chunks(ga, ga), chunks(gx, ga)])
    torch.testing.assert_close(sums, ref, rtol=1e-11, atol=1e-14)
    out, st = ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0)
    n_a = ref[1].sqrt()
    s = 500.0 / n_a
    torch.testing.assert_close(st[1].double(), n_a, rtol=1e-6, atol=0)
    torch.testing.assert_close(st[2].double(), s, rtol=1e-6, atol=0)
    # after the clip the gradient has norm min(1, ||x - s a||): here the NegGrad term has norm 500 -> clipped to 1
    out_norm = torch.sqrt(chunks(out, out))
    torch.testing.assert_close(out_norm, torch.ones_like(out_norm), rtol=1e-5, atol=0)
    torch.testing.assert_close(st[3].double() * st[4].double(), torch.ones((), dtype=torch.float64, device=dev), rtol=1e-5, atol=0)
    # idempotence of the clip: combining the already-clipped result with a zero NegGrad term leaves it unchanged
    sums2 = ops.norm3(out, out)
    out2, st2 = ops.combine(out, out, sums2, _lib.SISS_COMBINE_NONE, 0.0, 1.0)
    assert st2[4].item() >= 1.0 - 2e-6
    torch.testing.assert_close(out2, out, rtol=3e-6, atol=0)
    # linearity at fixed scalars: combine(2x, 2a) with scaling_norm doubled and no clip == 2 * (x - s a)
    o1, _ = ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 500.0, 0.0)
    s2 = ops.norm3(2 * gx, 2 * ga)
    o2, _ = ops.combine(2 * gx, 2 * ga, s2, _lib.SISS_COMBINE_SCALING_NORM, 1000.0, 0.0)
    torch.testing.assert_close(o2, 2 * o1, rtol=1e-6, atol=1e-12)


def test_pipelined_kernels_are_deterministic_and_match_eager_at_full_size(env):
    """Regression test for the TMA-ring release race (stage released before its LDS had completed):
    every pipelined kernel, run repeatedly at the full celeb shape, must reproduce itself bit for bit and
    match an eager evaluation on the device."""
    from siss_b200 import ops
    e = env
    go = 1 / 64
    noise_f = e["noise"].float()
    for it in range(24):
        a = ops.add_noise_pair(e["x0"], e["a0"], e["noise"], e["t"], e["ac"])
        f = ops.add_noise_mixture(e["x0"], e["a0"], e["noise"], e["keep"], e["t"], e["ac"], e["gamma"], e["sigma"], 0.5)
        k3 = ops.wmse_fwd_bwd(e["pred"], f[0], e["x0"], e["a0"], e["t"], e["gamma"], e["sigma"], f[3], f[4], go, go)
        dm = ops.dual_mse_fwd_bwd(e["pred"], e["pred"].flip(0).contiguous(), e["noise"], e["noise"], go, go)
        cur = list(a) + list(f) + list(k3) + list(dm)
        if it == 0:
            first = cur
            g32 = torch.tensor(go, device=e["dev"])
            assert torch.equal(dm[0], g32 * (2 * (e["pred"] - noise_f)))
            assert torch.equal(dm[1], g32 * (2 * (e["pred"].flip(0) - noise_f)))
            gm, sg = e["gamma"][e["t"]].view(-1, 1, 1, 1), e["sigma"][e["t"]].view(-1, 1, 1, 1)
            eps_a = (f[0].float() - gm * e["a0"].float()) / sg
            assert torch.equal(k3[1], (g32 * f[4]).view(-1, 1, 1, 1) * (2 * (e["pred"] - eps_a)))
        else:
            for x, y in zip(cur, first):
                assert torch.equal(x, y)


def test_k3_large_batch_takes_the_ldg_path_and_matches_eager(cuda_device):
    """B * D / W >= 4 M units selects the register-staged LDG kernel for K3 (measured faster there): check it
    against eager on the device at B = 192 x 3x256x256 bf16 (4.7 M units)."""
    from siss_b200 import ops
    from siss_b200.scheduler import SissDDPMScheduler
    dev = cuda_device
    Bl = 192
    g = torch.Generator(device=dev).manual_seed(5)
    shape = (Bl, C, H, W)
    sched = SissDDPMScheduler()
    gamma, sigma = sched.gamma_sigma(dev)
    x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
    a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
    xm = torch.randn(shape, device=dev, generator=g).bfloat16()
    pred = torch.randn(shape, device=dev, generator=g)
    t = torch.randint(200, 1000, (Bl,), device=dev, generator=g)
    w_x, w_a = torch.rand(Bl, device=dev, generator=g) * 2, torch.rand(Bl, device=dev, generator=g) * 2
    go = 1 / 64
    gx, ga, rlx, rla = ops.wmse_fwd_bwd(pred, xm, x0, a0, t, gamma, sigma, w_x, w_a, go, go)
    gm, sg = gamma[t].view(-1, 1, 1, 1), sigma[t].view(-1, 1, 1, 1)
    eps_x = (xm.float() - gm * x0.float()) / sg
    eps_a = (xm.float() - gm * a0.float()) / sg
    g32 = torch.tensor(go, device=dev)
    assert torch.equal(gx, (g32 * w_x).view(-1, 1, 1, 1) * (2 * (pred - eps_x)))
    assert torch.equal(ga, (g32 * w_a).view(-1, 1, 1, 1) * (2 * (pred - eps_a)))
    torch.testing.assert_close(rlx.double(), ((pred - eps_x).double() ** 2).sum(dim=[1, 2, 3]), rtol=2e-6, atol=0)
    gx2, ga2, _, _ = ops.wmse_fwd_bwd(pred, xm, x0, a0, t, gamma, sigma, w_x, w_a, go, go)
    assert torch.equal(gx, gx2) and torch.equal(ga, ga2)


def test_ldg_kernels_pass_the_parity_suite(cuda_device):
    """The plain-LDG variants of the row kernels (scalar path, K3 at large sizes, SISS_NO_TMA=1) are product
    code too: re-run the kernel parity tests in a subprocess with the TMA pipeline switched off — and through the
    ctypes binding (SISS_BINDING=ctypes), so that both bindings see the whole parity suite (this process uses the torch
    extension)."""
    import os
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    env = dict(os.environ, SISS_NO_TMA="1", SISS_BINDING="ctypes")
    cmd = [sys.executable, "-m", "pytest", str(root / "tests" / "test_kernels_gpu.py"), str(root / "tests" / "test_fuzz_gpu.py"),
           str(root / "tests" / "test_rng_gpu.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider"]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=str(root))
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-2000:]
    assert " passed" in out.stdout
