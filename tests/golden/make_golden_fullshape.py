"""Generate tests/golden/fullshape_*.npz: the REFERENCE's own SISS loss at the real CelebA-HQ shape
(3x256x256, D = 196 608), where the importance weights are hardest to pin — each Gaussian exponent is a
fp32 sum of ~98 304 (t = 999) and their DIFFERENCE is fed to exp().

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden_fullshape.py

A full-size case is ~20 MB of tensors per sample, far too much for a committed fixture, so the fixture holds
  * the seed the inputs are regenerated from (torch CPU generator: x0, a0 ~ U(-1,1), eps ~ N(0,1), exactly the
    statements of ``regenerate`` below) and a SHA-256 of every regenerated tensor, so a test can prove it rebuilt
    the very inputs the reference saw;
  * everything per-sample the reference returned: importance weights, per-row sums (float64) of loss_x, loss_a,
    weighted_loss_x, weighted_loss_a and of the autograd gradients into the UNet output;
  * a strided sample (every 4099th element) of loss_x, loss_a, weighted losses and both gradients for bit-exact
    element-wise checks;
  * the float64 evaluation of the same weights and the measured reference-vs-float64 deviation, which is what the
    test tolerance in tests/helpers.py::weight_tolerance is derived from.
"""
import hashlib
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
from make_golden import StubUNet, add_noise, alphas_cumprod, load_reference  # noqa: E402

STRIDE = 4099

CASES = [
    dict(name="fullshape_celeb_bf16_t999", B=4, dtype="bfloat16", t=[999, 999, 999, 999], lambd=0.5, seed=9001),
    dict(name="fullshape_celeb_fp32_tmix", B=4, dtype="float32", t=[999, 940, 860, 999], lambd=0.5, seed=9002),
]


def regenerate(seed: int, B: int, dtype: torch.dtype):
    """The inputs of a full-shape case, bit for bit (torch CPU generator). Tests call this too."""
    g = torch.Generator().manual_seed(seed)
    shape = (B, 3, 256, 256)
    x0 = (torch.rand(shape, generator=g) * 2 - 1).to(dtype)
    a0 = (torch.rand(shape, generator=g) * 2 - 1).to(dtype)
    noise = torch.randn(shape, generator=g).to(dtype)
    return x0, a0, noise


def sha(t: torch.Tensor) -> str:
    t = t.detach().contiguous()
    raw = t.view(torch.int16).numpy().tobytes() if t.dtype in (torch.bfloat16, torch.float16) else t.numpy().tobytes()
    return hashlib.sha256(raw).hexdigest()


def weights64(m, x0, a0, g, s, lambd):
    g, s = g.double(), s.double()
    m, x0, a0 = m.double(), x0.double(), a0.double()
    dx = ((m - g[:, None, None, None] * x0) ** 2).sum(dim=[1, 2, 3]) / (2 * s ** 2)
    da = ((m - g[:, None, None, None] * a0) ** 2).sum(dim=[1, 2, 3]) / (2 * s ** 2)
    return 1 / ((1 - lambd) + lambd * torch.exp(dx - da)), 1 / ((1 - lambd) * torch.exp(da - dx) + lambd), dx, da


def make_case(c, RefLoss):
    dt = getattr(torch, c["dtype"])
    B = c["B"]
    x0, a0, noise = regenerate(c["seed"], B, dt)
    t = torch.tensor(c["t"]).long()
    ac = alphas_cumprod("linear")
    gamma, sigma = ac ** 0.5, (1 - ac) ** 0.5
    xt_x, xt_a = add_noise(ac, x0, noise, t), add_noise(ac, a0, noise, t)
    all_d = {"og_latents": x0, "noisy_latents": xt_x}
    del_d = {"og_latents": a0, "noisy_latents": xt_a}
    unet = StubUNet()
    # make the reference's internal `torch.rand(B) > lambd` reproduce `keep`: replay the draw
    mask_seed = c["seed"] + 1000
    torch.manual_seed(mask_seed)
    items = RefLoss(gamma=gamma, sigma=sigma).importance_sampling_with_mixture(unet, t, noise, {}, all_d, del_d,
                                                                                lambd=c["lambd"])
    torch.manual_seed(mask_seed)
    keep = torch.rand(B) > c["lambd"]
    pred = unet.last_pred
    pred.grad = None
    (items[5].sum() / B).backward(retain_graph=True)
    gx = pred.grad.clone()
    pred.grad = None
    (items[6].sum() / B).backward()
    ga = pred.grad.clone()
    m = torch.where(keep[:, None, None, None], xt_x, xt_a)
    w64x, w64a, dx64, da64 = weights64(m, x0, a0, gamma[t], sigma[t], c["lambd"])
    out = dict(seed=np.array(c["seed"]), mask_seed=np.array(mask_seed), B=np.array(B), dtype=np.array(c["dtype"]),
               t=t.numpy(), lambd=np.array(c["lambd"]), keep_mask=keep.numpy(), stride=np.array(STRIDE),
               sha_x0=np.array(sha(x0)), sha_a0=np.array(sha(a0)), sha_noise=np.array(sha(noise)),
               sha_pred=np.array(sha(pred)), sha_x_mix=np.array(sha(m)),
               w_x=items[3].numpy(), w_a=items[4].numpy(),
               w_x_f64=w64x.numpy(), w_a_f64=w64a.numpy(), dist_x_f64=dx64.numpy(), dist_a_f64=da64.numpy())
    flat = lambda v: v.detach().reshape(B, -1)
    for k, v in (("loss_x", items[1]), ("loss_a", items[2]), ("wl_x", items[5]), ("wl_a", items[6]), ("grad_x", gx),
                 ("grad_a", ga)):
        out[f"rowsum_{k}"] = flat(v).double().sum(dim=1).numpy()
        out[f"sample_{k}"] = flat(v)[:, ::STRIDE].contiguous().numpy()
    rel = torch.maximum(((items[3].double() - w64x) / w64x).abs().max(), ((items[4].double() - w64a) / w64a).abs().max())
    out["ref_vs_f64_max_rel"] = np.array(float(rel))
    return out


def main():
    RefLoss = load_reference()
    for c in CASES:
        data = make_case(c, RefLoss)
        path = HERE / f"{c['name']}.npz"
        np.savez_compressed(path, **data)
        print(f"{path.name}: {path.stat().st_size / 1024:.1f} KiB  w_x={data['w_x']}  w_x_f64={data['w_x_f64']}  "
              f"reference-vs-float64 max rel dev {float(data['ref_vs_f64_max_rel']):.3e}")


if __name__ == "__main__":
    main()
