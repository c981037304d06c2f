"""CPU oracle for the SISS hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline. Nothing under
``siss_b200/`` imports it; the product path has no CPU fallback.

What it is: a restatement, in eager PyTorch on CPU tensors (the reference's own runtime, so dtype
promotion and bf16 rounding are the reference's), of

  * ``losses/ddpm_deletion_loss.py`` (all six methods)                      -> ``OracleDeletionLoss``
  * diffusers==0.27.2 ``DDPMScheduler`` betas / ``alphas_cumprod`` / ``add_noise``
    (third-party, pinned at environment.yml:232, NOT present under /root/reference and not
    installed here; restated from that release's published algorithm)     -> ``make_alphas_cumprod``, ``add_noise``
  * the inline two-backward + gradient-combine block, delete_celeb.py:682-767
    (tshirt inf-guard delete_tshirt.py:688-690)                           -> ``ReferenceGradLoop``, ``combine_flat``
  * ``accelerator.backward`` (loss / gradient_accumulation_steps, accelerate==0.27.2) and
    ``accelerator.clip_grad_norm_`` (= ``torch.nn.utils.clip_grad_norm_``, called directly).

Parity pinning:
  * loss class: PINNED — tests/golden/*.npz are outputs of the reference's own file executed from
    /root/reference by tests/golden/make_golden.py; tests/test_oracle_golden.py checks this module
    against them bit for bit. Where the reference checkout exists (the build container),
    tests/test_oracle_vs_reference_live.py also compares against the imported reference class on
    randomized cases (outputs and gradients, bit-exact).
  * combine block (``ReferenceGradLoop``): PINNED in the build container — the same live test reads the
    block's own source lines from delete_celeb.py / delete_tshirt.py / delete_sd.py, executes them with
    stub ``accelerator`` / ``cfg`` / ``wandb`` objects and compares gradients and logged scalars bit for
    bit; tests/golden/step_*.npz (whole optimiser steps produced by the reference's loss class and that block
    together, tests/golden/make_golden_step.py) carry the pin to the GPU box (tests/test_step_golden.py).
    It is also cross-checked against real autograd on a small module
    (tests/test_oracle_golden.py::test_combine_matches_literal_loop).
  * membership metric: PINNED (tests/golden/membership_*.npz from the reference class + the live test).
  * add_noise / beta schedules (diffusers), ``accelerator.backward`` scaling and ``clip_grad_norm_``
    routing (accelerate), EMAModel (diffusers): third-party packages that are neither vendored by the
    reference nor installed here — restated from their published behaviour, PARITY UNPINNED at that
    boundary (the reference holds no test, fixture or golden vector for them).

Every function is dtype-generic: call it with float64 tensors to get the "exact" answer the fp32
results are compared against.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------------
# Noise schedule and add_noise (diffusers 0.27.2 DDPMScheduler; call sites delete_celeb.py:602-603)
# --------------------------------------------------------------------------------------------------
def make_alphas_cumprod(num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                        beta_schedule: str = "linear") -> Tensor:
    """fp32 cumulative product of (1 - beta). ``linear`` is the DDPM/celebahq/tshirt schedule
    (config/train_tshirt_mnist.yaml:43-50), ``scaled_linear`` is Stable Diffusion's."""
    if beta_schedule == "linear":
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    elif beta_schedule == "scaled_linear":
        betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    else:
        raise NotImplementedError(beta_schedule)
    return torch.cumprod(1.0 - betas, dim=0)


def gamma_sigma(alphas_cumprod: Tensor) -> Tuple[Tensor, Tensor]:
    """delete_celeb.py:367-371: gamma = abar**0.5, sigma = (1 - abar)**0.5 (fp32 tables)."""
    return alphas_cumprod ** 0.5, (1 - alphas_cumprod) ** 0.5


def _ieee_sqrt(v: Tensor) -> Tensor:
    """``v ** 0.5`` with a correctly rounded square root. ATen evaluates ``pow(x, 0.5)`` as ``sqrt``
    in fp32 (16-bit inputs are widened first) and rounds to ``v.dtype``. On CUDA that sqrt is IEEE;
    torch-CPU's AVX-512 vectorised sqrt on the build host is NOT (about 0.6 % of inputs are 1 ulp off
    the correctly rounded value — measured), so the oracle takes the root in float64
    and rounds, which is exact for fp32. ``add_noise`` runs on the sample's device in the reference,
    i.e. on the GPU, so IEEE is the behaviour to restate."""
    if v.dtype == torch.float64:
        return v.sqrt()
    return v.double().sqrt().float().to(v.dtype)


def add_noise(alphas_cumprod: Tensor, original_samples: Tensor, noise: Tensor, timesteps: Tensor) -> Tensor:
    """DDPMScheduler.add_noise: the table is cast to the SAMPLE dtype first, then gathered, then
    square-rooted; the two products and the sum are ordinary eager ops in the sample dtype."""
    ac = alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
    timesteps = timesteps.to(original_samples.device)
    sqrt_alpha_prod = _ieee_sqrt(ac[timesteps]).flatten()
    while sqrt_alpha_prod.dim() < original_samples.dim():
        sqrt_alpha_prod = sqrt_alpha_prod.unsqueeze(-1)
    sqrt_one_minus = _ieee_sqrt(1 - ac[timesteps]).flatten()
    while sqrt_one_minus.dim() < original_samples.dim():
        sqrt_one_minus = sqrt_one_minus.unsqueeze(-1)
    return sqrt_alpha_prod * original_samples + sqrt_one_minus * noise


# --------------------------------------------------------------------------------------------------
# Loss class (losses/ddpm_deletion_loss.py)
# --------------------------------------------------------------------------------------------------
def _bcast(v: Tensor, like: Tensor) -> Tensor:
    return v.reshape(v.shape[0], *([1] * (like.dim() - 1)))


def draw_keep_mask(batch_size: int, lambd: float) -> Tensor:
    """ddpm_deletion_loss.py:18 — CPU default-generator uniform draw; True = keep-batch row."""
    return torch.rand(batch_size) > lambd


def select_mixture(noisy_keep: Tensor, noisy_forget: Tensor, keep_mask: Tensor) -> Tensor:
    """ddpm_deletion_loss.py:19-23 — row select."""
    mix = torch.empty_like(noisy_keep)
    forget_mask = ~keep_mask
    mix[keep_mask] = noisy_keep[keep_mask]
    mix[forget_mask] = noisy_forget[forget_mask]
    return mix


def gaussian_exponents(mix: Tensor, x0: Tensor, a0: Tensor, gamma_t: Tensor, sigma_t: Tensor) -> Tuple[Tensor, Tensor]:
    """ddpm_deletion_loss.py:31-39 — d = ||x_t - gamma x0||^2 / (2 sigma^2), summed over all but dim 0."""
    dims = list(range(1, mix.dim()))
    g = _bcast(gamma_t, mix)
    d_x = torch.sum((mix - g * x0) ** 2, dim=dims)
    d_x = d_x / (2 * (sigma_t ** 2))
    d_a = torch.sum((mix - g * a0) ** 2, dim=dims)
    d_a = d_a / (2 * (sigma_t ** 2))
    return d_x, d_a


def importance_weights(d_x: Tensor, d_a: Tensor, lambd: float) -> Tuple[Tensor, Tensor]:
    """ddpm_deletion_loss.py:41-45."""
    r_ax = torch.exp(d_x - d_a)
    r_xa = torch.exp(d_a - d_x)
    w_x = 1 / ((1 - lambd) + lambd * r_ax)
    w_a = 1 / ((1 - lambd) * r_xa + lambd)
    return w_x, w_a


class OracleDeletionLoss:
    """Restatement of DDPMDeletionLoss; same constructor, method names, parameters and 7-tuples.
    ``keep_mask`` is an extra optional argument so tests can share one draw between oracle and kernel."""

    def __init__(self, gamma: Tensor, sigma: Tensor):
        self.all_gamma = gamma
        self.all_sigma = sigma

    def importance_sampling_with_mixture(self, unet, timesteps, noise, conditioning, all_samples_dict,
                                         deletion_samples_dict, lambd, keep_mask: Optional[Tensor] = None):
        g_t = self.all_gamma[timesteps]
        s_t = self.all_sigma[timesteps]
        noisy_keep, noisy_forget = all_samples_dict['noisy_latents'], deletion_samples_dict['noisy_latents']
        x0, a0 = all_samples_dict['og_latents'], deletion_samples_dict['og_latents']
        if keep_mask is None:
            keep_mask = draw_keep_mask(noisy_keep.shape[0], lambd)
        mix = select_mixture(noisy_keep, noisy_forget, keep_mask)
        pred = unet(mix, timesteps, **conditioning, return_dict=False)[0]

        eps_x = (mix - _bcast(g_t, mix) * x0) / _bcast(s_t, mix)     # :26
        eps_a = (mix - _bcast(g_t, mix) * a0) / _bcast(s_t, mix)     # :27
        loss_x = (pred - eps_x) ** 2                                  # :29
        loss_a = (pred - eps_a) ** 2                                  # :30
        d_x, d_a = gaussian_exponents(mix, x0, a0, g_t, s_t)          # :32-39
        w_x, w_a = importance_weights(d_x, d_a, lambd)                # :41-45
        wl_x = _bcast(w_x, loss_x) * loss_x                           # :51
        wl_a = _bcast(w_a, loss_a) * loss_a                           # :53
        return None, loss_x, loss_a, w_x, w_a, wl_x, wl_a

    def double_forward_with_neg_del(self, unet, timesteps, noise, conditioning, all_samples_dict,
                                    deletion_samples_dict):
        pred_x = unet(all_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_x = (pred_x - noise) ** 2
        pred_a = unet(deletion_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_a = (pred_a - noise) ** 2
        return None, loss_x, loss_a, None, None, loss_x, loss_a

    def erasediff(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict,
                  uniform_noise: Optional[Tensor] = None):
        pred_x = unet(all_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_x = (pred_x - noise) ** 2
        pred_a = unet(deletion_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        if uniform_noise is None:
            uniform_noise = torch.rand_like(pred_a)                   # :75, drawn after forward #2
        loss_a = (pred_a - uniform_noise) ** 2
        return None, loss_x, loss_a, None, None, loss_x, loss_a

    def simple_neg_del(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict,
                       superfactor):
        pred_a = unet(deletion_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_a = (pred_a - noise) ** 2
        loss = -superfactor * loss_a
        return loss, None, loss_a, None, None, None, None

    def naive_del(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict):
        pred_x = unet(all_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_x = (pred_x - noise) ** 2
        return loss_x, loss_x, None, None, None, None, None

    def subscore_bernoulli(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict,
                           lambd, keep_mask: Optional[Tensor] = None):
        noisy_keep, noisy_forget = all_samples_dict['noisy_latents'], deletion_samples_dict['noisy_latents']
        if keep_mask is None:
            keep_mask = draw_keep_mask(noisy_keep.shape[0], lambd)
        forget_mask = ~keep_mask
        mix = select_mixture(noisy_keep, noisy_forget, keep_mask)
        pred = unet(mix, timesteps, **conditioning, return_dict=False)[0]
        loss = (pred - noise) ** 2
        loss_x = (1 / (1 - lambd)) * loss[keep_mask]
        loss_a = loss[forget_mask]
        if len(loss_x) == 0:                                          # :113-116
            loss_x = torch.zeros(1, 1, 1, 1, requires_grad=True)
            loss_a = torch.zeros(1, 1, 1, 1, requires_grad=True)
        if len(loss_a) == 0:                                          # :118-120
            loss_a = torch.zeros(1, 1, 1, 1, requires_grad=True)
        return None, loss_x, loss_a, None, None, loss_x, loss_a


# --------------------------------------------------------------------------------------------------
# Two backward passes + gradient combine (delete_celeb.py:682-767)
# --------------------------------------------------------------------------------------------------
def accelerate_backward(loss: Tensor, grad_accum_steps: int, retain_graph: bool = False) -> None:
    """accelerate==0.27.2 ``Accelerator.backward`` without a scaler: divide by the accumulation
    steps, then ``.backward()``."""
    (loss / grad_accum_steps).backward(retain_graph=retain_graph)


class ReferenceGradLoop:
    """The reference's inline gradient bookkeeping for one optimiser step, restated around an
    ``nn.Module``: per-parameter dicts keyed by name, clones and subtractions included, so the op
    and rounding order is the reference's. Use: ``micro_step`` G times, then ``sync_step``."""

    def __init__(self, model: torch.nn.Module, train_batch_size: int, grad_accum_steps: int = 1):
        self.model = model
        self.train_batch_size = train_batch_size
        self.grad_accum_steps = grad_accum_steps
        self.accum_loss_x: Dict[str, Tensor] = {}
        self.accum_loss_a: Dict[str, Tensor] = {}

    def micro_step(self, items: Sequence[Optional[Tensor]], retain_graph: bool) -> None:
        loss, _lx, _la, _wx, _wa, weighted_loss_x, weighted_loss_a = items
        if loss is not None:                                                          # :682-684
            accelerate_backward(loss.sum() / self.train_batch_size, self.grad_accum_steps)
            return
        wl_x = weighted_loss_x.sum() / self.train_batch_size                          # :686
        wl_a = weighted_loss_a.sum() / self.train_batch_size                          # :687
        accelerate_backward(wl_x, self.grad_accum_steps, retain_graph=retain_graph)   # :691
        grads_x = {n: p.grad.clone() for n, p in self.model.named_parameters()}       # :693-696
        accelerate_backward(wl_a, self.grad_accum_steps)                              # :702
        for n, p in self.model.named_parameters():                                    # :705-711
            true_grad = p.grad.clone() - grads_x[n]
            if n not in self.accum_loss_a:
                self.accum_loss_a[n] = true_grad
            else:
                self.accum_loss_a[n] += true_grad

    def sync_step(self, single_loss: bool, loss_fn: str = "importance_sampling_with_mixture",
                  scaling_norm: Optional[float] = None, eta: Optional[float] = None, max_norm: float = 1.0,
                  inf_guard: bool = False) -> Dict[str, Tensor]:
        out: Dict[str, Tensor] = {}
        if not single_loss:
            for n, p in self.model.named_parameters():                                # :717-718
                self.accum_loss_x[n] = p.grad.clone() - self.accum_loss_a[n]
            sq_x = 0.0
            sq_a = 0.0
            for g in self.accum_loss_x.values():                                      # :725-726
                sq_x = sq_x + torch.norm(g, p=2) ** 2
            for g in self.accum_loss_a.values():                                      # :729-730
                sq_a = sq_a + torch.norm(g, p=2) ** 2
            norm_x = torch.sqrt(sq_x)                                                 # :733-734
            norm_a = torch.sqrt(sq_a)
            if loss_fn == "erasediff":                                                # :740-742
                dot = sum(torch.sum(self.accum_loss_x[n] * self.accum_loss_a[n])
                          for n, _ in self.model.named_parameters())
                scaling_factor = eta - dot / (norm_a ** 2)
                scaling_factor = -max(scaling_factor, 0)
            else:                                                                     # :746
                scaling_factor = scaling_norm / norm_a
                if inf_guard and torch.isinf(scaling_factor):                         # delete_tshirt.py:688-690
                    scaling_factor = 0
            for n, p in self.model.named_parameters():                                # :749-750
                p.grad = self.accum_loss_x[n] - scaling_factor * self.accum_loss_a[n]
            out.update(norm_x=norm_x, norm_a=norm_a, scaling_factor=torch.as_tensor(scaling_factor))
            self.accum_loss_x, self.accum_loss_a = {}, {}
        if max_norm is not None:
            out["total_norm"] = torch.nn.utils.clip_grad_norm_(self.model.parameters(), max_norm)  # :767
        return out


def combine_flat(g_x: Tensor, g_a: Tensor, scaling_norm: Optional[float] = None, eta: Optional[float] = None,
                 max_norm: Optional[float] = 1.0, inf_guard: bool = False, loss_scale: Optional[float] = None):
    """The sync-step arithmetic of ``ReferenceGradLoop.sync_step`` on ONE flat tensor pair (what the
    combine amounts to when the model has a single parameter tensor). Returns
    (grad, norm_x, norm_a, scaling_factor, total_norm, clip_coef).

    ``loss_scale`` restates the ``mixed_precision: fp16`` run (delete_celeb.py:104,237-242; accelerate==0.27.2 and
    torch's GradScaler, environment.yml:222 — third party, parity unpinned): ``g_x`` / ``g_a`` are gradients of the
    SCALED loss; the norms, the scaling factor and the combination (:725-750) are taken on them as they are, then
    ``accelerator.clip_grad_norm_`` first unscales (``GradScaler.unscale_``: grad *= 1/scale) and then clips (:767)."""
    norm_x = torch.sqrt(torch.norm(g_x, p=2) ** 2)
    norm_a = torch.sqrt(torch.norm(g_a, p=2) ** 2)
    if eta is not None:
        s = eta - torch.sum(g_x * g_a) / (norm_a ** 2)
        s = -max(s, 0)
    elif scaling_norm is not None:
        s = scaling_norm / norm_a
        if inf_guard and torch.isinf(s):
            s = 0
    else:
        s = 0
    grad = g_x - s * g_a
    if loss_scale is not None:
        grad = grad * (1.0 / float(loss_scale))
    total_norm = torch.norm(grad, p=2)
    clip = torch.ones((), dtype=grad.dtype)
    if max_norm is not None:
        clip = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
        grad = grad * clip
    return grad, norm_x, norm_a, torch.as_tensor(s, dtype=grad.dtype), total_norm, clip


def norm3_cpu(g_x: Tensor, g_a: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """CPU stand-in for K4a (``siss_norm3``): the three sums in float64. Used ONLY by the gloo tests
    that exercise GradCombiner's collective choreography on a CPU box."""
    r = torch.stack([(g_x.double() ** 2).sum(), (g_a.double() ** 2).sum(), (g_x.double() * g_a.double()).sum()])
    if out is None:
        return r
    out.copy_(r)
    return out


def combine_from_sums_cpu(g_x: Tensor, g_a: Tensor, sums3: Tensor, mode: int, value: float, max_norm: float = 1.0,
                          inf_guard: bool = False, out: Optional[Tensor] = None, stats: Optional[Tensor] = None):
    """CPU stand-in for K4b (``siss_combine``): same scalar logic (scaling factor per mode, norm of the
    combination obtained algebraically from the three sums, clip coefficient), fp32 element-wise
    update. mode: 0 scaling_norm, 1 erasediff, 2 none. Test infrastructure only."""
    f32 = torch.float32
    sxx, saa, sxa = (float(v) for v in sums3.tolist())
    n_x = torch.sqrt(torch.tensor(sxx, dtype=f32))
    n_a = torch.sqrt(torch.tensor(saa, dtype=f32))
    if mode == 2:
        s = torch.zeros((), dtype=f32)
    elif mode == 1:
        s = torch.tensor(value, dtype=f32) - torch.tensor(sxa, dtype=f32) / (n_a * n_a)
        s = -(s if not (0.0 > s) else torch.zeros((), dtype=f32))
    else:
        s = torch.tensor(value, dtype=f32) / n_a
        if inf_guard and torch.isinf(s):
            s = torch.zeros((), dtype=f32)
    sd = float(s)
    tn = torch.tensor(max(sxx - 2.0 * sd * sxa + sd * sd * saa, 0.0), dtype=torch.float64).sqrt().to(f32)
    clip = torch.ones((), dtype=f32)
    if max_norm and max_norm > 0:
        clip = torch.clamp(torch.tensor(max_norm, dtype=f32) / (tn + 1e-6), max=1.0)
    res = (g_x - s * g_a) * clip
    if out is None:
        out = res
    else:
        out.copy_(res)
    if stats is not None:
        stats.copy_(torch.stack([n_x, n_a, s, tn, clip]))
    return out, stats


# --------------------------------------------------------------------------------------------------
# EMA of the parameters (delete_celeb.py:776-777 -> diffusers.training_utils.EMAModel, diffusers==0.27.2,
# NOT vendored by the reference: restated from that release, parity unpinned at this third-party boundary).
# Config keys: ema.use_ema / ema_max_decay / ema_inv_gamma / ema_power (config/train_tshirt_mnist.yaml:94-97).
# --------------------------------------------------------------------------------------------------
class OracleEMA:
    def __init__(self, parameters, decay: float = 0.9999, min_decay: float = 0.0, update_after_step: int = 0,
                 use_ema_warmup: bool = False, inv_gamma: float = 1.0, power: float = 2.0 / 3.0):
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.decay, self.min_decay, self.update_after_step = decay, min_decay, update_after_step
        self.use_ema_warmup, self.inv_gamma, self.power = use_ema_warmup, inv_gamma, power
        self.optimization_step = 0
        self.cur_decay_value = None

    def get_decay(self, optimization_step: int) -> float:
        step = max(0, optimization_step - self.update_after_step - 1)
        if step <= 0:
            return 0.0
        if self.use_ema_warmup:
            cur = 1 - (1 + step / self.inv_gamma) ** -self.power
        else:
            cur = (1 + step) / (10 + step)
        cur = min(cur, self.decay)
        return max(cur, self.min_decay)

    @torch.no_grad()
    def step(self, parameters) -> None:
        self.optimization_step += 1
        decay = self.get_decay(self.optimization_step)
        self.cur_decay_value = decay
        one_minus_decay = 1 - decay
        for s_param, param in zip(self.shadow_params, parameters):
            if param.requires_grad:
                s_param.sub_(one_minus_decay * (s_param - param))
            else:
                s_param.copy_(param)


# --------------------------------------------------------------------------------------------------
# Membership-loss metric (metrics/class_membership.py:69-128), pinned by tests/golden/membership_*.npz which were
# produced by executing the reference class itself (tests/golden/make_golden_membership.py).
# --------------------------------------------------------------------------------------------------
@torch.no_grad()
def membership_losses(all_images: Tensor, deletion_images: Tensor, noise: Tensor, alphas_cumprod: Tensor, unet,
                      timesteps: Sequence[int], eval_batch_size: int) -> List[List[Tensor]]:
    n_img, n_noise = all_images.shape[0], noise.shape[0]
    out = []
    for timestep in timesteps:
        # :76-86 — image i repeated over the noise draws, the noise set tiled over the images
        all_flat = all_images.unsqueeze(1).expand(-1, n_noise, -1, -1, -1).reshape(-1, *all_images.shape[1:])
        del_flat = deletion_images.unsqueeze(1).expand(-1, n_noise, -1, -1, -1).reshape(-1, *all_images.shape[1:])
        noise_flat = noise.unsqueeze(0).expand(n_img, -1, -1, -1, -1).reshape(-1, *noise.shape[1:])
        t_flat = torch.full((all_flat.shape[0],), timestep)                        # :89
        noisy_all = add_noise(alphas_cumprod, all_flat, noise_flat, t_flat)        # :92-93
        noisy_del = add_noise(alphas_cumprod, del_flat, noise_flat, t_flat)
        all_l, del_l = [], []
        for i in range(0, all_flat.shape[0], eval_batch_size):                     # :100-112
            nz = noise_flat[i:i + eval_batch_size]
            ts = t_flat[:nz.shape[0]]
            p_all = unet(noisy_all[i:i + eval_batch_size], ts, return_dict=False)[0]
            p_del = unet(noisy_del[i:i + eval_batch_size], ts, return_dict=False)[0]
            all_l.append(torch.sum((p_all - nz) ** 2, dim=[1, 2, 3]))
            del_l.append(torch.sum((p_del - nz) ** 2, dim=[1, 2, 3]))
        out.append([torch.mean(torch.cat(all_l)), torch.mean(torch.cat(del_l))])   # :114-115
    return out


# --------------------------------------------------------------------------------------------------
# Per-batch statistics (delete_celeb.py:626-656) — consumed by the stats tests
# --------------------------------------------------------------------------------------------------
def batch_stats(items: Sequence[Optional[Tensor]]) -> Dict[str, float]:
    loss, loss_x, loss_a, w_x, w_a, _wlx, _wla = items
    stats: Dict[str, float] = {}
    for name, t in (("loss", loss), ("loss_x", loss_x), ("loss_a", loss_a)):
        if t is not None:
            per = t.mean(dim=list(range(1, t.dim())))
            stats[f"{name}/mean"] = t.mean().item()
            stats[f"{name}/max"] = per.max().item()
            stats[f"{name}/min"] = per.min().item()
            stats[f"{name}/std"] = per.std().item()
    for name, t in (("importance_weight_x", w_x), ("importance_weight_a", w_a)):
        if t is not None:
            stats[f"{name}/mean"] = t.mean().item()
            stats[f"{name}/max"] = t.max().item()
            stats[f"{name}/min"] = t.min().item()
            stats[f"{name}/std"] = t.std().item()
    return stats


class StubUNet(torch.nn.Module):
    """UNet stand-in for oracle/kernel parity and for timing the path without the UNet masking it
    (BASELINE.md §4): one scale parameter and one bias, honouring the call convention
    ``unet(x, t, **conditioning, return_dict=False)[0]`` (ddpm_deletion_loss.py:24)."""

    def __init__(self, init_scale: float = 0.75, init_bias: float = 0.05, dtype=torch.float32):
        super().__init__()
        self.scale = torch.nn.Parameter(torch.tensor(init_scale, dtype=dtype))
        self.bias = torch.nn.Parameter(torch.tensor(init_bias, dtype=dtype))

    def forward(self, x, timesteps, encoder_hidden_states=None, return_dict=False, **kw):
        out = x.to(self.scale.dtype) * self.scale + self.bias
        if encoder_hidden_states is not None:
            out = out + encoder_hidden_states.to(out.dtype).mean() * 0.01
        return (out,)
