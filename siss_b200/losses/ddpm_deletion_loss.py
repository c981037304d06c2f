"""``DDPMDeletionLoss`` — drop-in for the reference's losses/ddpm_deletion_loss.py.

Same constructor (``gamma``, ``sigma``), same method names, same positional/keyword parameters
(the SD task calls by keyword, delete_sd.py:977-985, so names are part of the contract) and the
same 7-tuple ``(loss, loss_x, loss_a, importance_weight_x, importance_weight_a, weighted_loss_x,
weighted_loss_a)``. The returned tensors are autograd-connected to the UNet exactly like the
reference's, so ``.sum() / B`` + ``backward(retain_graph=...)`` in the task loops run unchanged;
all ``[B, C, H, W]`` arithmetic is done by the sm_100a kernels through autograd Functions.

The fast path that never materialises the four loss tensors is ``siss_b200.step.UnlearnStep``.
"""
from __future__ import annotations

import torch

from .. import ops


class _WeightedSqErr(torch.autograd.Function):
    """(pred) -> loss_x, loss_a, w_x*loss_x, w_a*loss_a   (ddpm_deletion_loss.py:26-30, 51-53)."""

    @staticmethod
    def forward(ctx, pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a)
        return ops.wmse_fwd(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a)

    @staticmethod
    def backward(ctx, g_lx, g_la, g_wx, g_wa):
        pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a = ctx.saved_tensors
        grad = ops.wmse_bwd(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a, g_lx, g_la, g_wx, g_wa)
        return (grad,) + (None,) * 8


class _SqErr(torch.autograd.Function):
    """(pred, target) -> (pred - target)**2"""

    @staticmethod
    def forward(ctx, pred, target):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pred, target)
        return ops.sqerr_fwd(pred, target)[0]

    @staticmethod
    def backward(ctx, g):
        pred, target = ctx.saved_tensors
        if g is None:
            return None, None
        return ops.sqerr_bwd(pred, target, go_loss=g), None


class _SqErrScaled(torch.autograd.Function):
    """(pred, target, alpha) -> loss, alpha*loss   (simple_neg_del: loss = -superfactor * loss_a)."""

    @staticmethod
    def forward(ctx, pred, target, alpha):
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(pred, target)
        ctx.alpha = float(alpha)
        return ops.sqerr_fwd(pred, target, alpha=ctx.alpha)

    @staticmethod
    def backward(ctx, g_loss, g_scaled):
        pred, target = ctx.saved_tensors
        if g_loss is None and g_scaled is None:
            return None, None, None
        return ops.sqerr_bwd(pred, target, go_loss=g_loss, go_scaled=g_scaled, alpha=ctx.alpha), None, None


def _draw_keep_mask(batch_size: int, lambd: float) -> torch.Tensor:
    """The reference's Bernoulli draw, verbatim in kind and position: a CPU ``torch.rand`` from the
    default generator (ddpm_deletion_loss.py:18), so identical seeds give identical masks."""
    return torch.rand(batch_size) > lambd


class DDPMDeletionLoss:
    def __init__(self, gamma, sigma):
        # gamma = sqrt(alphas_cumprod), sigma = sqrt(1 - alphas_cumprod): 1-D fp32 tables
        self.all_gamma = gamma
        self.all_sigma = sigma

    # ---- SISS ------------------------------------------------------------------------------------
    def importance_sampling_with_mixture(self, unet, timesteps, noise, conditioning, all_samples_dict,
                                         deletion_samples_dict, lambd, keep_mask=None):
        """ddpm_deletion_loss.py:11-56. ``noise`` is accepted and unused, as in the reference.
        ``keep_mask`` (optional, bool[B]) overrides the CPU Bernoulli draw — used for data-parallel
        sharding, where every rank slices one global draw."""
        noisy_x = all_samples_dict['noisy_latents']
        noisy_a = deletion_samples_dict['noisy_latents']
        x0 = all_samples_dict['og_latents']
        a0 = deletion_samples_dict['og_latents']
        batch_size = noisy_x.shape[0]
        all_mask = _draw_keep_mask(batch_size, lambd) if keep_mask is None else keep_mask

        x_mix, _dist_x, _dist_a, w_x, w_a = ops.mixture_weights(
            noisy_x, noisy_a, x0, a0, all_mask, timesteps, self.all_gamma, self.all_sigma, lambd)
        mixture_noise_preds = unet(x_mix, timesteps, **conditioning, return_dict=False)[0]
        loss_x, loss_a, weighted_loss_x, weighted_loss_a = _WeightedSqErr.apply(
            mixture_noise_preds, x_mix, x0, a0, timesteps, self.all_gamma, self.all_sigma, w_x, w_a)
        return None, loss_x, loss_a, w_x, w_a, weighted_loss_x, weighted_loss_a

    # ---- SISS (No IS) ------------------------------------------------------------------------------
    def double_forward_with_neg_del(self, unet, timesteps, noise, conditioning, all_samples_dict,
                                    deletion_samples_dict):
        """ddpm_deletion_loss.py:60-67."""
        all_noise_preds = unet(all_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_x = _SqErr.apply(all_noise_preds, noise)
        deletion_noise_preds = unet(deletion_samples_dict['noisy_latents'], timesteps, **conditioning,
                                    return_dict=False)[0]
        loss_a = _SqErr.apply(deletion_noise_preds, noise)
        return None, loss_x, loss_a, None, None, loss_x, loss_a

    # ---- EraseDiff ---------------------------------------------------------------------------------
    def erasediff(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict):
        """ddpm_deletion_loss.py:70-78. The uniform target is drawn with the same torch call at the
        same point (after the second forward), so the device RNG stream matches the reference's."""
        all_noise_preds = unet(all_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_x = _SqErr.apply(all_noise_preds, noise)
        deletion_noise_preds = unet(deletion_samples_dict['noisy_latents'], timesteps, **conditioning,
                                    return_dict=False)[0]
        uniform_noise = torch.rand_like(deletion_noise_preds)
        loss_a = _SqErr.apply(deletion_noise_preds, uniform_noise)
        return None, loss_x, loss_a, None, None, loss_x, loss_a

    # ---- NegGrad -----------------------------------------------------------------------------------
    def simple_neg_del(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict,
                       superfactor):
        """ddpm_deletion_loss.py:82-88."""
        deletion_noise_preds = unet(deletion_samples_dict['noisy_latents'], timesteps, **conditioning,
                                    return_dict=False)[0]
        loss_a, loss = _SqErrScaled.apply(deletion_noise_preds, noise, -superfactor)
        return loss, None, loss_a, None, None, None, None

    # ---- Naive deletion ------------------------------------------------------------------------------
    def naive_del(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict):
        """ddpm_deletion_loss.py:91-96."""
        all_noise_preds = unet(all_samples_dict['noisy_latents'], timesteps, **conditioning, return_dict=False)[0]
        loss_x = _SqErr.apply(all_noise_preds, noise)
        loss = loss_x
        return loss, loss_x, None, None, None, None, None

    # ---- Reviewer-proposed loss (only the tshirt task honours it, delete_tshirt.py:632) -------------
    def subscore_bernoulli(self, unet, timesteps, noise, conditioning, all_samples_dict, deletion_samples_dict,
                           lambd, keep_mask=None):
        """ddpm_deletion_loss.py:99-122. Same Bernoulli row select as SISS (K2's select), one forward,
        plain squared error against the shared noise (the sqerr kernel), then the reference's ragged split
        by mask with the 1/(1-lambd) factor on the keep rows. The boolean-mask gathers are torch indexing:
        the outputs are ragged, there is nothing to fuse. lambd == 1 raises ZeroDivisionError exactly like
        the reference (Python-level 1 / (1 - lambd))."""
        noisy_x = all_samples_dict['noisy_latents']
        noisy_a = deletion_samples_dict['noisy_latents']
        batch_size = noisy_x.shape[0]
        all_mask = _draw_keep_mask(batch_size, lambd) if keep_mask is None else keep_mask
        deletion_mask = ~all_mask
        bernoulli_samples, *_ = ops.mixture_weights(
            noisy_x, noisy_a, all_samples_dict['og_latents'], deletion_samples_dict['og_latents'], all_mask,
            timesteps, self.all_gamma, self.all_sigma, lambd)
        noise_preds = unet(bernoulli_samples, timesteps, **conditioning, return_dict=False)[0]
        loss = _SqErr.apply(noise_preds, noise)
        dev_all, dev_del = all_mask.to(loss.device), deletion_mask.to(loss.device)
        loss_x = (1 / (1 - lambd)) * loss[dev_all]
        loss_a = loss[dev_del]
        if len(loss_x) == 0:
            print('no nondeletion samples')
            loss_x = torch.zeros(1, 1, 1, 1, requires_grad=True, device=loss.device)
            loss_a = torch.zeros(1, 1, 1, 1, requires_grad=True, device=loss.device)
        if len(loss_a) == 0:
            print('no deletion samples')
            loss_a = torch.zeros(1, 1, 1, 1, requires_grad=True, device=loss.device)
        return None, loss_x, loss_a, None, None, loss_x, loss_a
