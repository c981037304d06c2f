"""The SISS loss at the REAL CelebA-HQ shape (3x256x256, D = 196 608) against outputs of the reference's own class.

`tests/golden/fullshape_*.npz` (made by `tests/golden/make_golden_fullshape.py`, which executes
/root/reference/losses/ddpm_deletion_loss.py) hold the seed the inputs are regenerated from, SHA-256 digests that
prove the regeneration, and the reference's per-sample outputs / strided element samples.

Tolerances (DESIGN.md §3 "Importance weights"):
  * element-wise losses, weighted losses and gradients (fed the reference's own weights): bit-exact;
  * per-row sums: rtol 1e-5 against the float64 sums of the reference's tensors;
  * importance weights vs the reference's fp32 weights: rtol = 1.8 eps32 (d_x + d_a) + 1e-5 — 1.8 is < 2x the
    largest deviation of the REFERENCE ITSELF from the float64 evaluation measured at this shape (0.93 in those
    units, B = 64, fp32 / bf16, t = 999 / uniform); at t = 999 that is ~4 % instead of the 19 % window of round 1;
  * importance weights vs the float64 evaluation: rtol 2e-4 (measured 1e-5-level: the kernel accumulates the
    exponent DIFFERENCE directly instead of subtracting two ~1e5-sized fp32 sums).
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

from helpers import FULLSHAPE_CASES, GOLDEN_DIR, weight_tolerance
from oracle import siss_oracle as O

sys.path.insert(0, str(GOLDEN_DIR))
from make_golden_fullshape import regenerate, sha  # noqa: E402


def _load(name):
    z = np.load(GOLDEN_DIR / f"{name}.npz", allow_pickle=False)
    c = {k: z[k] for k in z.files}
    dt = getattr(torch, str(c["dtype"]))
    B = int(c["B"])
    x0, a0, noise = regenerate(int(c["seed"]), B, dt)
    assert sha(x0) == str(c["sha_x0"]) and sha(a0) == str(c["sha_a0"]) and sha(noise) == str(c["sha_noise"]), \
        "regenerated inputs differ from the ones the reference saw (torch CPU generator changed?)"
    return c, dt, B, x0, a0, noise, torch.from_numpy(c["t"]).long(), torch.from_numpy(c["keep_mask"])


def test_fixtures_exist():
    assert len(FULLSHAPE_CASES) >= 2


@pytest.mark.parametrize("name", FULLSHAPE_CASES)
def test_oracle_matches_reference_at_full_shape(name):
    """CPU: the oracle's restatement reproduces the reference's outputs at D = 196 608 bit for bit."""
    c, dt, B, x0, a0, noise, t, keep = _load(name)
    ac = O.make_alphas_cumprod()
    gamma, sigma = O.gamma_sigma(ac)
    all_d = {"og_latents": x0, "noisy_latents": O.add_noise(ac, x0, noise, t)}
    del_d = {"og_latents": a0, "noisy_latents": O.add_noise(ac, a0, noise, t)}
    unet = O.StubUNet()
    items = O.OracleDeletionLoss(gamma, sigma).importance_sampling_with_mixture(
        unet, t, noise, {}, all_d, del_d, lambd=float(c["lambd"]), keep_mask=keep)
    assert np.array_equal(items[3].numpy(), c["w_x"]) and np.array_equal(items[4].numpy(), c["w_a"])
    st = int(c["stride"])
    for k, v in (("loss_x", items[1]), ("loss_a", items[2]), ("wl_x", items[5]), ("wl_a", items[6])):
        assert np.array_equal(v.detach().reshape(B, -1)[:, ::st].numpy(), c[f"sample_{k}"]), k
        np.testing.assert_allclose(v.detach().reshape(B, -1).double().sum(dim=1).numpy(), c[f"rowsum_{k}"], rtol=1e-12)
    # the recorded deviation of the reference from float64 is what the weight tolerance is derived from
    eps = np.finfo(np.float32).eps
    dsum = c["dist_x_f64"] + c["dist_a_f64"]
    for w, w64 in ((c["w_x"], c["w_x_f64"]), (c["w_a"], c["w_a_f64"])):
        rel = np.abs(w.astype(np.float64) - w64) / w64
        assert (rel <= 0.93 * eps * dsum + 1e-5).all(), "the reference deviates more than recorded in DESIGN.md"


@pytest.mark.gpu
@pytest.mark.parametrize("name", FULLSHAPE_CASES)
def test_kernels_match_reference_at_full_shape(name, cuda_device):
    from siss_b200 import ops
    from siss_b200.scheduler import SissDDPMScheduler
    dev = cuda_device
    c, dt, B, x0, a0, noise, t, keep = _load(name)
    sched = SissDDPMScheduler()
    ac = sched.alphas_cumprod.to(dev)
    gamma, sigma = sched.gamma_sigma(dev)
    lam = float(c["lambd"])
    X0, A0, NZ, T = x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev)
    x_mix, d_x, d_a, w_x, w_a = ops.add_noise_mixture(X0, A0, NZ, keep, T, ac, gamma, sigma, lam)
    assert sha(x_mix.cpu()) == str(c["sha_x_mix"]), "x_mix differs from the reference's mixture sample"
    # weights: vs the reference (tolerance derived from the reference's own error) and vs float64
    tol = weight_tolerance(d_x.cpu(), d_a.cpu())
    for got, ref, f64, nm in ((w_x, c["w_x"], c["w_x_f64"], "w_x"), (w_a, c["w_a"], c["w_a_f64"], "w_a")):
        got = got.cpu().double()
        rel_ref = (got - torch.from_numpy(ref).double()).abs() / torch.from_numpy(ref).double()
        assert (rel_ref <= tol).all(), f"{nm}: vs reference {rel_ref} > {tol}"
        rel64 = (got - torch.from_numpy(f64)).abs() / torch.from_numpy(f64)
        assert (rel64 <= 2e-4).all(), f"{nm}: vs float64 {rel64}"
    torch.testing.assert_close(d_x.cpu().double(), torch.from_numpy(c["dist_x_f64"]), rtol=1e-5, atol=0)
    torch.testing.assert_close(d_a.cpu().double(), torch.from_numpy(c["dist_a_f64"]), rtol=1e-5, atol=0)

    pred = x_mix.float() * 0.75 + 0.05                     # the fixture's StubUNet (make_golden.py)
    assert sha(pred.cpu()) == str(c["sha_pred"])
    rw_x, rw_a = torch.from_numpy(c["w_x"]).to(dev), torch.from_numpy(c["w_a"]).to(dev)   # the reference's weights
    st = int(c["stride"])
    outs = ops.wmse_fwd(pred, x_mix, X0, A0, T, gamma, sigma, rw_x, rw_a)
    for k, v in zip(("loss_x", "loss_a", "wl_x", "wl_a"), outs):
        v = v.reshape(B, -1)
        assert np.array_equal(v[:, ::st].cpu().numpy(), c[f"sample_{k}"]), f"{k}: sampled elements not bit-exact"
        np.testing.assert_allclose(v.double().sum(dim=1).cpu().numpy(), c[f"rowsum_{k}"], rtol=1e-12)
    go = float(np.float32(1.0) / np.float32(B))
    g_x, g_a, rl_x, rl_a = ops.wmse_fwd_bwd(pred, x_mix, X0, A0, T, gamma, sigma, rw_x, rw_a, go, go)
    for k, v in (("grad_x", g_x), ("grad_a", g_a)):
        v = v.reshape(B, -1)
        assert np.array_equal(v[:, ::st].cpu().numpy(), c[f"sample_{k}"]), f"{k}: sampled elements not bit-exact"
        np.testing.assert_allclose(v.double().sum(dim=1).cpu().numpy(), c[f"rowsum_{k}"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(rl_x.cpu().double().numpy(), c["rowsum_loss_x"], rtol=1e-5)
    np.testing.assert_allclose(rl_a.cpu().double().numpy(), c["rowsum_loss_a"], rtol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("tmode", ["t999", "uniform"])
def test_weight_deviation_at_headline_batch(dtype, tmode, cuda_device):
    """B = 64 x 3x256x256 (the bench shape): the kernel's weights against (i) the reference formula evaluated by eager
    torch ON THE DEVICE in the reference's op order (losses/ddpm_deletion_loss.py:32-45) and (ii) float64. The measured
    deviations are appended to gpurun_out/weights_pin.jsonl (copied to profiles/ and quoted in DESIGN.md)."""
    import json
    from siss_b200 import ops
    from siss_b200.scheduler import SissDDPMScheduler
    dev = cuda_device
    B, shape, lam = 64, (64, 3, 256, 256), 0.5
    g = torch.Generator(device=dev).manual_seed(321)
    x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dtype)
    a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dtype)
    nz = torch.randn(shape, device=dev, generator=g).to(dtype)
    t = torch.full((B,), 999, device=dev) if tmode == "t999" else torch.randint(0, 1000, (B,), device=dev, generator=g)
    keep = torch.rand(B, device=dev, generator=g) > lam
    sched = SissDDPMScheduler()
    ac = sched.alphas_cumprod.to(dev)
    gamma, sigma = sched.gamma_sigma(dev)
    x_mix, d_x, d_a, w_x, w_a = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, lam)

    def weights(dt):
        gg, ss = gamma[t].to(dt), sigma[t].to(dt)
        m, x, a = (x_mix, x0, a0) if dt == torch.float32 else (x_mix.double(), x0.double(), a0.double())
        dx = ((m - gg[:, None, None, None] * x) ** 2).sum(dim=[1, 2, 3])
        dx /= (2 * (ss ** 2))
        da = ((m - gg[:, None, None, None] * a) ** 2).sum(dim=[1, 2, 3])
        da /= (2 * (ss ** 2))
        return 1 / ((1 - lam) + lam * torch.exp(dx - da)), 1 / ((1 - lam) * torch.exp(da - dx) + lam), dx, da

    rx, ra, _, _ = weights(torch.float32)
    fx, fa, dx64, da64 = weights(torch.float64)
    live = (fx > 1e-30) & (fa > 1e-30) & torch.isfinite(fx) & torch.isfinite(fa)
    assert int(live.sum()) >= 8

    def rel(p, q):
        return float(torch.maximum(((p.double() - q.double()) / q.double()).abs()[live].max(), torch.tensor(0.0, device=dev)))

    eps = torch.finfo(torch.float32).eps
    unit = (eps * (dx64 + da64))[live]
    k_ref = float(torch.maximum((((rx.double() - fx) / fx).abs()[live] / unit).max(),
                                (((ra.double() - fa) / fa).abs()[live] / unit).max()))
    rec = {"dtype": str(dtype), "timesteps": tmode, "B": B, "live_rows": int(live.sum()),
           "kernel_vs_eager_ref_max_rel": max(rel(w_x, rx), rel(w_a, ra)),
           "eager_ref_vs_f64_max_rel": max(rel(rx, fx), rel(ra, fa)),
           "kernel_vs_f64_max_rel": max(rel(w_x, fx), rel(w_a, fa)),
           "eager_ref_vs_f64_in_eps32_times_dsum": k_ref}
    out = Path(__file__).resolve().parent.parent / "gpurun_out"
    out.mkdir(exist_ok=True)
    with open(out / "weights_pin.jsonl", "a") as f:
        f.write(json.dumps(rec) + "\n")
    tol = weight_tolerance(d_x, d_a).to(dev)
    for got, ref, f64 in ((w_x, rx, fx), (w_a, ra, fa)):
        r_ref = ((got.double() - ref.double()) / ref.double()).abs()
        assert (r_ref[live] <= tol[live]).all(), rec
        assert ((got.double() - f64) / f64).abs()[live].max() <= 2e-4, rec
