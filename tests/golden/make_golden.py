"""Generate tests/golden/*.npz by executing the REFERENCE's own loss class.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

Each fixture holds the seeded inputs and the outputs of
``/root/reference/losses/ddpm_deletion_loss.py::DDPMDeletionLoss`` for every method, plus the
autograd gradients of ``weighted_loss.sum() / B`` into the UNet output. The noise schedule tables and
the noisy latents are produced with plain torch expressions here (diffusers is not installed), and are
stored in the fixture, so the loss-class outputs are pinned for exactly those inputs.
bf16 tensors are stored widened to fp32 (lossless) with their dtype recorded.
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/losses/ddpm_deletion_loss.py")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_ddpm_deletion_loss", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DDPMDeletionLoss


class StubUNet(torch.nn.Module):
    """1-scale/1-bias stand-in, fp32 output (what accelerate's autocast wrapper returns)."""

    def __init__(self):
        super().__init__()
        self.scale = torch.nn.Parameter(torch.tensor(0.75))
        self.bias = torch.nn.Parameter(torch.tensor(0.05))
        self.last_pred = None

    def forward(self, x, timesteps, encoder_hidden_states=None, return_dict=False, **kw):
        out = x.to(torch.float32) * self.scale + self.bias
        if encoder_hidden_states is not None:
            out = out + encoder_hidden_states.to(out.dtype).mean() * 0.01
        out.retain_grad()
        self.last_pred = out
        return (out,)


def alphas_cumprod(schedule):
    if schedule == "linear":
        betas = torch.linspace(1e-4, 0.02, 1000, dtype=torch.float32)
    else:
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


def add_noise(ac, x0, noise, t):
    a = ac.to(x0.dtype)
    sa = (a[t] ** 0.5).reshape(-1, 1, 1, 1)
    s1 = ((1 - a[t]) ** 0.5).reshape(-1, 1, 1, 1)
    return sa * x0 + s1 * noise


def npy(t):
    if t is None:
        return None
    t = t.detach()
    if t.dtype in (torch.bfloat16, torch.float16):
        t = t.float()
    return t.numpy()


CASES = [
    # name, B, C, H, W, dtype, schedule, timesteps spec, lambd, cond
    dict(name="tshirt_fp32", B=8, C=1, H=28, W=28, dtype="float32", schedule="linear", t="uniform", lambd=0.5),
    dict(name="celeb_bf16_t999", B=4, C=3, H=16, W=16, dtype="bfloat16", schedule="linear", t=999, lambd=0.5),
    dict(name="celeb_fp32_t999", B=6, C=3, H=16, W=16, dtype="float32", schedule="linear", t=999, lambd=0.5),
    dict(name="sd_fp32_cond", B=2, C=4, H=8, W=8, dtype="float32", schedule="scaled_linear", t=999, lambd=0.5,
         cond=True),
    dict(name="lambd0", B=5, C=1, H=6, W=6, dtype="float32", schedule="linear", t="uniform", lambd=0.0),
    dict(name="lambd1", B=5, C=1, H=6, W=6, dtype="float32", schedule="linear", t="uniform", lambd=1.0),
    dict(name="small_t_saturate", B=6, C=1, H=7, W=5, dtype="float32", schedule="linear", t="small", lambd=0.3),
    dict(name="fp16_mid_t", B=4, C=2, H=8, W=8, dtype="float16", schedule="linear", t="mid", lambd=0.5),
    dict(name="odd_D_fp32", B=3, C=1, H=5, W=3, dtype="float32", schedule="linear", t="uniform", lambd=0.5),
]


def make_case(c, RefLoss, seed):
    torch.manual_seed(seed)
    dt = getattr(torch, c["dtype"])
    shape = (c["B"], c["C"], c["H"], c["W"])
    x0 = (torch.rand(shape) * 2 - 1).to(dt)
    a0 = (torch.rand(shape) * 2 - 1).to(dt)
    noise = torch.randn(shape).to(dt)
    if c["t"] == "uniform":
        t = torch.randint(0, 1000, (c["B"],)).long()
    elif c["t"] == "small":
        t = torch.tensor([0, 1, 2, 5, 20, 60][: c["B"]]).long()
    elif c["t"] == "mid":
        t = torch.randint(600, 1000, (c["B"],)).long()
    else:
        t = torch.full((c["B"],), int(c["t"])).long()
    ac = alphas_cumprod(c["schedule"])
    gamma, sigma = ac ** 0.5, (1 - ac) ** 0.5
    xt_x, xt_a = add_noise(ac, x0, noise, t), add_noise(ac, a0, noise, t)
    cond = {}
    ehs = None
    if c.get("cond"):
        ehs = torch.randn(c["B"], 77, 16)
        cond = {"encoder_hidden_states": ehs}
    all_d = {"og_latents": x0, "noisy_latents": xt_x}
    del_d = {"og_latents": a0, "noisy_latents": xt_a}
    loss_obj = RefLoss(gamma=gamma, sigma=sigma)
    B = c["B"]
    out = dict(x0=npy(x0), a0=npy(a0), noise=npy(noise), t=t.numpy(), alphas_cumprod=ac.numpy(),
               gamma=gamma.numpy(), sigma=sigma.numpy(), xt_x=npy(xt_x), xt_a=npy(xt_a),
               dtype=np.array(c["dtype"]), lambd=np.array(c["lambd"]), seed=np.array(seed),
               schedule=np.array(c["schedule"]))
    if ehs is not None:
        out["encoder_hidden_states"] = ehs.numpy()

    def grads_into_pred(unet, scalar, retain):
        unet.last_pred.grad = None
        scalar.backward(retain_graph=retain)
        return unet.last_pred.grad.clone()

    # --- SISS -------------------------------------------------------------------------------------
    unet = StubUNet()
    mask_seed = seed + 1000
    torch.manual_seed(mask_seed)
    items = loss_obj.importance_sampling_with_mixture(unet, t, noise, cond, all_d, del_d, lambd=c["lambd"])
    torch.manual_seed(mask_seed)
    keep = torch.rand(B) > c["lambd"]          # replay of the reference's draw
    out["siss_mask_seed"] = np.array(mask_seed)
    out["siss_keep_mask"] = keep.numpy()
    out["siss_pred"] = npy(unet.last_pred)
    for k, v in zip(("loss_x", "loss_a", "w_x", "w_a", "wl_x", "wl_a"), items[1:]):
        out[f"siss_{k}"] = npy(v)
    gx = grads_into_pred(unet, items[5].sum() / B, True)
    ga = grads_into_pred(unet, items[6].sum() / B, False)
    out["siss_grad_x"], out["siss_grad_a"] = npy(gx), npy(ga)

    # --- No-IS ------------------------------------------------------------------------------------
    unet = StubUNet()
    items = loss_obj.double_forward_with_neg_del(unet, t, noise, cond, all_d, del_d)
    out["nois_loss_x"], out["nois_loss_a"] = npy(items[1]), npy(items[2])

    # --- EraseDiff --------------------------------------------------------------------------------
    unet = StubUNet()
    torch.manual_seed(seed + 2000)
    items = loss_obj.erasediff(unet, t, noise, cond, all_d, del_d)
    out["erasediff_seed"] = np.array(seed + 2000)
    out["erasediff_loss_x"], out["erasediff_loss_a"] = npy(items[1]), npy(items[2])

    # --- NegGrad / naive --------------------------------------------------------------------------
    unet = StubUNet()
    items = loss_obj.simple_neg_del(unet, t, noise, cond, all_d, del_d, superfactor=1.7)
    out["neg_loss"], out["neg_loss_a"] = npy(items[0]), npy(items[2])
    unet = StubUNet()
    items = loss_obj.naive_del(unet, t, noise, cond, all_d, del_d)
    out["naive_loss"] = npy(items[0])

    # --- subscore_bernoulli -----------------------------------------------------------------------
    unet = StubUNet()
    torch.manual_seed(seed + 3000)
    out["subscore_seed"] = np.array(seed + 3000)
    try:
        items = loss_obj.subscore_bernoulli(unet, t, noise, cond, all_d, del_d, lambd=c["lambd"])
        out["subscore_loss_x"], out["subscore_loss_a"] = npy(items[1]), npy(items[2])
    except ZeroDivisionError:  # the reference divides by (1 - lambd) in Python: lambd = 1 raises
        out["subscore_raises"] = np.array("ZeroDivisionError")
    return out


def main():
    if not REF.exists():
        sys.exit(f"{REF} not found: golden vectors can only be regenerated where the reference is mounted")
    RefLoss = load_reference()
    for i, c in enumerate(CASES):
        data = make_case(c, RefLoss, seed=1234 + 17 * i)
        path = HERE / f"{c['name']}.npz"
        np.savez_compressed(path, **{k: v for k, v in data.items() if v is not None})
        print(f"{path.name}: {path.stat().st_size / 1024:.1f} KiB, w_x={data['siss_w_x']}, w_a={data['siss_w_a']}")


if __name__ == "__main__":
    main()
