"""Fused fast path for one unlearning micro-step / optimiser step.

The drop-in class (:mod:`siss_b200.losses`) has to materialise the four ``[B,C,H,W]`` loss tensors
the reference's 7-tuple promises. A caller that owns its training loop does not need them: per
micro-step this driver launches

    K1oK2  siss_add_noise_mixture   x0, a0, eps -> x_mix, w_x, w_a                 (4 s B/elem)
           unet(x_mix, t, **conditioning)                                          (diffusers / any callable)
    K3     siss_wmse_fwd_bwd        eps_hat, x_mix, x0, a0 -> dL_x/deps_hat, dL_a/deps_hat, row sums
    autograd.backward(eps_hat, dL_x) into G_x ; autograd.backward(eps_hat, dL_a) into G_a

and per optimiser step K4a + K4b (:class:`siss_b200.grad_combine.GradCombiner`). It reproduces the
reference loop delete_celeb.py:580-767 (noise is drawn by the caller, as there) including the
``/ train_batch_size`` and ``/ gradient_accumulation_steps`` scalings, and returns only device tensors:
nothing here synchronises the host (the reference does ~20 ``.item()`` per micro-step, :626-663).
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch

from . import ops
from .grad_combine import GradCombiner
from .losses.ddpm_deletion_loss import _draw_keep_mask
from .scheduler import SissDDPMScheduler

TWO_TERM = ("importance_sampling_with_mixture", "double_forward_with_neg_del", "erasediff", "subscore_bernoulli")
ONE_TERM = ("naive_del", "simple_neg_del")


def _dual_mse(px: torch.Tensor, pa: torch.Tensor, tgt_x: torch.Tensor, tgt_a: torch.Tensor, go_x: float, go_a: float):
    """``ops.dual_mse_fwd_bwd`` with eager's type promotion: a 16-bit UNet output against an fp32 target is computed in
    fp32 and the gradients are rounded back to the prediction dtype, as autograd does (rare: accelerate's autocast
    wrapper returns fp32 predictions); a 16-bit target of another 16-bit kind than the prediction is widened to fp32."""
    if tgt_x.dtype != tgt_a.dtype:
        wide = torch.promote_types(tgt_x.dtype, tgt_a.dtype)
        tgt_x, tgt_a = tgt_x.to(wide), tgt_a.to(wide)
    if px.dtype != torch.float32 and tgt_x.dtype != px.dtype:
        if tgt_x.dtype != torch.float32:
            tgt_x, tgt_a = tgt_x.float(), tgt_a.float()
        g_x, g_a, rl_x, rl_a = ops.dual_mse_fwd_bwd(px.float(), pa.float(), tgt_x, tgt_a, go_x, go_a)
        return g_x.to(px.dtype), g_a.to(pa.dtype), rl_x, rl_a
    return ops.dual_mse_fwd_bwd(px, pa, tgt_x, tgt_a, go_x, go_a)


def upstream_scale(train_batch_size: int, grad_accum_steps: int) -> float:
    """d(total)/d(weighted_loss element) exactly as autograd forms it for
    ``(w.sum() / train_batch_size / G).backward()``: fp32(fp32(1/G) / train_batch_size)."""
    g = np.float32(1.0) / np.float32(grad_accum_steps)
    return float(np.float32(g) / np.float32(train_batch_size))


class UnlearnStep:
    """One rank's unlearning step. ``train_batch_size`` is the GLOBAL per-micro-step batch the loss is
    normalised by (``cfg.train_batch_size`` in the reference; under data parallel every rank passes its
    shard of the batch and the same global value)."""

    def __init__(self, unet: Callable, scheduler: SissDDPMScheduler, combiner: GradCombiner, *,
                 loss_fn: str = "importance_sampling_with_mixture", train_batch_size: int,
                 gradient_accumulation_steps: int = 1, lambd: Optional[float] = None,
                 superfactor: Optional[float] = None, scaling_norm: Optional[float] = None,
                 eta: Optional[float] = None, max_norm: Optional[float] = 1.0, inf_guard: bool = False,
                 superfactor_decay: Optional[float] = None, device_rng=None,
                 t_range: Optional[Tuple[int, int]] = None, superfactor_decay_on: str = "micro_step"):
        """``device_rng`` (a :class:`siss_b200.rng.DeviceRng`, opt-in): draw eps, the timesteps and the Bernoulli mask
        on the device from the counter-based stream whenever ``micro_step`` is not given them — for SISS eps is then
        generated inside K1oK2 and never touches HBM. ``t_range`` = [lo, hi) for drawn timesteps (default: the whole
        schedule, delete_tshirt.py:535-540; (999, 1000) reproduces delete_celeb.py:593-598)."""
        if loss_fn not in TWO_TERM + ONE_TERM:
            raise ValueError(f"unknown loss_fn {loss_fn!r}")
        if superfactor_decay_on not in ("micro_step", "sync_step"):
            raise ValueError('superfactor_decay_on must be "micro_step" or "sync_step"')
        if loss_fn in ("importance_sampling_with_mixture", "subscore_bernoulli") and lambd is None:
            raise ValueError(f"{loss_fn} needs lambd")
        if loss_fn == "simple_neg_del" and superfactor is None:
            raise ValueError("simple_neg_del needs superfactor")
        if loss_fn == "erasediff" and eta is None:
            raise ValueError("erasediff needs eta")
        if loss_fn in ("importance_sampling_with_mixture", "double_forward_with_neg_del", "subscore_bernoulli") \
                and scaling_norm is None:
            raise ValueError(f"{loss_fn} needs scaling_norm")
        self.unet, self.scheduler, self.combiner = unet, scheduler, combiner
        self.loss_fn = loss_fn
        self.train_batch_size = int(train_batch_size)
        self.G = int(gradient_accumulation_steps)
        self.lambd, self.superfactor = lambd, superfactor
        # `deletion.superfactor_decay`: the reference multiplies loss_params.superfactor by it after every
        # micro-step's statistics (delete_celeb.py:658-662), i.e. the NEXT micro-step sees the decayed value
        self.superfactor_decay = superfactor_decay
        # ... in delete_celeb.py / delete_tshirt.py (:600-604). delete_sd.py applies the decay once per OPTIMISER step
        # instead, under `if accelerator.sync_gradients:` (delete_sd.py:1173-1193): superfactor_decay_on="sync_step".
        self.superfactor_decay_on = superfactor_decay_on
        self.scaling_norm, self.eta, self.max_norm, self.inf_guard = scaling_norm, eta, max_norm, inf_guard
        self.go = upstream_scale(self.train_batch_size, self.G)
        dev = combiner.device
        self.alphas_cumprod = scheduler.alphas_cumprod.to(dev)
        self.gamma, self.sigma = scheduler.gamma_sigma(dev)
        self._micro = 0
        self.device_rng = device_rng
        # optional profiling callback: called with a stage name on the current stream right after each stage of a
        # micro-step has been enqueued ("k1k2", "unet_fwd", "k3", "backward_x", "backward_a"); bench.py records CUDA
        # events in it to decompose the end-to-end step. None (the default) costs nothing.
        self.stage_hook: Optional[Callable[[str], None]] = None
        self.t_range = (0, int(self.gamma.numel())) if t_range is None else (int(t_range[0]), int(t_range[1]))

    # ------------------------------------------------------------------------------------------
    def micro_step(self, x0: torch.Tensor, a0: torch.Tensor, noise: Optional[torch.Tensor] = None,
                   timesteps: Optional[torch.Tensor] = None,
                   conditioning: Optional[dict] = None, keep_mask: Optional[torch.Tensor] = None,
                   forget_target: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """Forward + both backward passes for one micro-batch. Returns per-sample device tensors:
        ``row_loss_x`` / ``row_loss_a`` (sum over C,H,W of the unweighted squared errors) and, for SISS,
        ``w_x`` / ``w_a`` / ``dist_x`` / ``dist_a``. ``forget_target`` (EraseDiff only) replaces the uniform draw
        of ddpm_deletion_loss.py:75 with a given tensor — for reproducible tests and replays."""
        cond = conditioning or {}
        out: Dict[str, torch.Tensor] = {}
        last = self._micro == self.G - 1          # last accumulation micro-step of this optimiser step
        siss = self.loss_fn == "importance_sampling_with_mixture"
        noise, timesteps, keep_mask, draw = self._device_draws(x0, noise, timesteps, keep_mask, siss, out)
        if siss:
            self._micro_siss(x0, a0, noise, timesteps, keep_mask, cond, draw, last, out)
        elif self.loss_fn == "subscore_bernoulli":
            self._micro_subscore(x0, a0, noise, timesteps, keep_mask, cond, last, out)
        elif self.loss_fn in ("double_forward_with_neg_del", "erasediff"):
            draw = self._micro_two_forward(x0, a0, noise, timesteps, forget_target, cond, draw, last, out)
        else:
            self._micro_single_term(x0, a0, noise, timesteps, cond, out)
        if self.superfactor is not None and self.superfactor_decay is not None and self.superfactor_decay_on == "micro_step":
            self.superfactor *= self.superfactor_decay
        if draw is not None:
            self.device_rng.advance()      # next micro-step draws from the next index (host mirror + device counter)
        self._micro += 1
        return out

    def _stage(self, name: str) -> None:
        if self.stage_hook is not None:
            self.stage_hook(name)

    def _device_draws(self, x0, noise, timesteps, keep_mask, siss: bool, out: Dict[str, torch.Tensor]):
        """Opt-in device RNG: fill in whatever of (timesteps, keep mask, eps) the caller left out. For SISS eps stays
        None — it is generated inside K1oK2. Returns (noise, timesteps, keep_mask, draw index or None)."""
        if noise is not None and timesteps is not None:
            return noise, timesteps, keep_mask, None
        rng = self.device_rng
        if rng is None:
            raise ValueError("noise= and timesteps= are required unless the step was built with device_rng=")
        draw = rng.next_draw()             # the calls below pass no draw=: that would bypass the device-side counter
        if timesteps is None or (siss and keep_mask is None):
            ts_d, keep_d = rng.draw_rows(x0.shape[0], x0.device, t_range=self.t_range if timesteps is None else None,
                                         lambd=self.lambd if (siss and keep_mask is None) else None)
            timesteps = ts_d if timesteps is None else timesteps
            keep_mask = keep_d if keep_d is not None else keep_mask
        if noise is None and not siss:
            noise = rng.randn(x0.shape, x0.dtype, x0.device)
        out["timesteps"] = timesteps
        return noise, timesteps, keep_mask, draw

    def _micro_siss(self, x0, a0, noise, timesteps, keep_mask, cond, draw, last: bool, out) -> None:
        """importance_sampling_with_mixture: K1oK2 -> UNet -> K3 -> two backward passes from one forward."""
        cb, rng = self.combiner, self.device_rng
        keep = _draw_keep_mask(x0.shape[0], self.lambd) if keep_mask is None else keep_mask
        if noise is None:     # eps generated inside K1oK2 (registers only)
            per_row = x0.numel() // max(x0.shape[0], 1)
            x_mix, d_x, d_a, w_x, w_a, _ = ops.add_noise_mixture_rng(
                x0, a0, keep, timesteps, self.alphas_cumprod, self.gamma, self.sigma, self.lambd, rng.seed, draw,
                elem_offset=rng.row_offset * per_row, d_draw=rng.d_draw)
        else:
            x_mix, d_x, d_a, w_x, w_a = ops.add_noise_mixture(x0, a0, noise, keep, timesteps, self.alphas_cumprod,
                                                              self.gamma, self.sigma, self.lambd)
        self._stage("k1k2")
        pred = self.unet(x_mix, timesteps, **cond, return_dict=False)[0]
        self._stage("unet_fwd")
        g_x, g_a, rl_x, rl_a = ops.wmse_fwd_bwd(pred.detach(), x_mix, x0, a0, timesteps, self.gamma, self.sigma,
                                                w_x, w_a, self.go, self.go)
        self._stage("k3")
        cb.begin_x()
        torch.autograd.backward(pred, g_x, retain_graph=True)
        self._stage("backward_x")
        cb.begin_a(last_micro_step=last)      # data parallel: G_x's reduce-scatter overlaps backward #2
        torch.autograd.backward(pred, g_a)
        self._stage("backward_a")
        out.update(w_x=w_x, w_a=w_a, dist_x=d_x, dist_a=d_a, row_loss_x=rl_x, row_loss_a=rl_a)

    def _micro_subscore(self, x0, a0, noise, timesteps, keep_mask, cond, last: bool, out) -> None:
        """subscore_bernoulli (ddpm_deletion_loss.py:99-122; only delete_tshirt.py runs it through the two-backward block,
        with retain_graph, :632): Bernoulli row select of the two noisy batches (the select of K1oK2; its weights are not
        used), ONE forward, squared error against the shared eps; the keep rows — scaled by the Python float 1/(1-lambd)
        — form the first term, the forget rows the second. In autograd's order that is, per element,
            keep row:    dL_x = fp32(go * fp32(1/(1-lambd))) * (2 u),   dL_a = 0
            forget row:  dL_x = 0,                                     dL_a = go * (2 u),      u = eps_hat - eps
        which the dual-MSE kernel evaluates for all rows with the two scalars; the row masks are applied to its outputs.
        The reference's zero-row fallbacks (:113-120) replace BOTH terms by constants when no keep row was drawn and the
        second term when no forget row was drawn — i.e. zero gradients, reproduced here on the device without a sync."""
        cb = self.combiner
        coef = 1 / (1 - self.lambd)        # ZeroDivisionError at lambd == 1, exactly like the reference (:111)
        keep = _draw_keep_mask(x0.shape[0], self.lambd) if keep_mask is None else keep_mask
        x_sel, *_ = ops.add_noise_mixture(x0, a0, noise, keep, timesteps, self.alphas_cumprod, self.gamma, self.sigma,
                                          self.lambd)
        pred = self.unet(x_sel, timesteps, **cond, return_dict=False)[0]
        go_x = float(np.float32(self.go) * np.float32(coef))
        p = pred.detach()
        g_x, g_a, rl, _ = _dual_mse(p, p, noise, noise, go_x, self.go)
        m = keep.to(device=p.device, dtype=torch.bool).view(-1, *([1] * (p.dim() - 1)))
        any_keep = m.any()
        zero = torch.zeros((), dtype=g_x.dtype, device=p.device)
        g_x = torch.where(m, g_x, zero)
        g_a = torch.where(m | ~any_keep, zero, g_a)
        cb.begin_x()
        torch.autograd.backward(pred, g_x, retain_graph=True)
        cb.begin_a(last_micro_step=last)
        torch.autograd.backward(pred, g_a)
        out.update(row_loss=rl, keep_mask=m.view(-1))

    def _micro_two_forward(self, x0, a0, noise, timesteps, forget_target, cond, draw, last: bool, out):
        """double_forward_with_neg_del / erasediff: K1 pair -> two UNet forwards -> dual-MSE kernel -> two backward passes.
        Returns the draw index used (EraseDiff with the device RNG takes one even when eps and t were given)."""
        cb, rng = self.combiner, self.device_rng
        xt_x, xt_a = self.scheduler.add_noise_pair(x0, a0, noise, timesteps)
        pred_x = self.unet(xt_x, timesteps, **cond, return_dict=False)[0]
        pred_a = self.unet(xt_a, timesteps, **cond, return_dict=False)[0]
        # EraseDiff's forget target: uniform noise drawn after the second forward (ddpm_deletion_loss.py:75)
        if self.loss_fn == "erasediff" and forget_target is None and rng is not None:
            # opt-in device RNG: the uniform target is drawn inside the kernel (aux domain of this micro-step's draw)
            if draw is None:
                draw = rng.next_draw()
            per_row = x0.numel() // max(x0.shape[0], 1)
            g_x, g_a, rl_x, rl_a, _ = ops.dual_mse_rng_fwd_bwd(
                pred_x.detach(), pred_a.detach(), noise, self.go, self.go, rng.seed, draw,
                elem_offset=rng.row_offset * per_row, d_draw=rng.d_draw)
        else:
            if self.loss_fn == "erasediff":
                tgt_a = torch.rand_like(pred_a) if forget_target is None else forget_target
            else:
                tgt_a = noise
            tgt_x = noise
            if tgt_a.dtype != tgt_x.dtype:
                # eager promotes (pred - noise) and (pred - uniform) independently (ddpm_deletion_loss.py:62,72-76); the
                # kernel takes one target dtype, so both go to the WIDER one — never round fp32 noise down to 16 bits
                wide = torch.promote_types(tgt_x.dtype, tgt_a.dtype)
                tgt_x, tgt_a = tgt_x.to(wide), tgt_a.to(wide)
            g_x, g_a, rl_x, rl_a = _dual_mse(pred_x.detach(), pred_a.detach(), tgt_x, tgt_a, self.go, self.go)
        cb.begin_x()
        torch.autograd.backward(pred_x, g_x)
        cb.begin_a(last_micro_step=last)
        torch.autograd.backward(pred_a, g_a)
        out.update(row_loss_x=rl_x, row_loss_a=rl_a)
        return draw

    def _micro_single_term(self, x0, a0, noise, timesteps, cond, out) -> None:
        """naive_del / simple_neg_del: one forward, one backward, no combine (delete_celeb.py:682-684)."""
        if self.loss_fn == "naive_del":
            xt = self.scheduler.add_noise(x0, noise, timesteps)
            alpha = 1.0
        else:
            xt = self.scheduler.add_noise(a0, noise, timesteps)
            alpha = -float(self.superfactor)
        pred = self.unet(xt, timesteps, **cond, return_dict=False)[0]
        # grad = (go * alpha) * 2 (pred - eps): dual kernel with the second term switched off
        g, _unused, rl, _ = _dual_mse(pred.detach(), pred.detach(), noise, noise,
                                      float(np.float32(self.go) * np.float32(alpha)), 0.0)
        self.combiner.begin_x()
        torch.autograd.backward(pred, g)
        out["row_loss_a" if self.loss_fn == "simple_neg_del" else "row_loss_x"] = rl

    def end_of_optimizer_step(self) -> None:
        """Bookkeeping at the accumulation boundary: restart the micro-step count and, for
        ``superfactor_decay_on="sync_step"``, apply the decay (delete_sd.py:1190-1193). ``sync_step`` calls it; call it
        yourself when the boundary is handled by ``FusedCombineAdamW.step`` instead."""
        self._micro = 0
        if self.superfactor is not None and self.superfactor_decay is not None and self.superfactor_decay_on == "sync_step":
            self.superfactor *= self.superfactor_decay

    def sync_step(self) -> torch.Tensor:
        """Gradient combine + clip at the accumulation boundary (delete_celeb.py:714-767). Leaves
        ``param.grad`` ready for ``optimizer.step()``; returns the device stats tensor
        ``[norm_loss_x, norm_loss_a, scaling_factor, total_norm, clip_coef]``."""
        self.end_of_optimizer_step()
        if self.loss_fn in ONE_TERM:
            return self.combiner.clip_only(self.max_norm if self.max_norm is not None else 0.0)
        if self.loss_fn == "erasediff":
            return self.combiner.combine(eta=self.eta, max_norm=self.max_norm, inf_guard=self.inf_guard)
        return self.combiner.combine(scaling_norm=self.scaling_norm, max_norm=self.max_norm,
                                     inf_guard=self.inf_guard)

    @property
    def is_sync_step(self) -> bool:
        return self._micro >= self.G


def batch_stats(out: Dict[str, torch.Tensor], elems_per_sample: int,
                dest: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The reference's per-batch statistics (delete_celeb.py:626-656) from the O(B) row sums of
    ``micro_step`` — ONE kernel launch (``siss_batch_stats``), 16 fp32 scalars on the device in the order
    of ``siss_b200.ops.STAT_KEYS``; no host synchronisation (the reference does ~20 ``.item()`` here)."""
    return ops.batch_stats(out.get("row_loss_x"), out.get("row_loss_a"), out.get("w_x"), out.get("w_a"),
                           elems_per_sample, out=dest)


class StepLog:
    """Sync-free logging: one small pinned-host record per step.

    ``push(stats16, grad_stats5)`` enqueues a non-blocking device->host copy of the 21 scalars into a ring
    of pinned buffers and records an event; ``pop_ready()`` returns the records whose copies have
    completed, as dicts keyed like the reference's ``batch_stats`` / wandb scalars — the training loop
    never waits for the GPU in order to log."""

    GRAD_KEYS = ("gradient/norm_loss_x", "gradient/norm_loss_a", "gradient/scaling_factor", "gradient/total_norm",
                 "gradient/clip_coef")

    def __init__(self, depth: int = 8):
        self._bufs = [torch.empty(21, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self._events = [torch.cuda.Event() for _ in range(depth)]
        self._pending = []   # (slot, step)
        self._next = 0

    def push(self, stats16: Optional[torch.Tensor], grad_stats5: Optional[torch.Tensor], step: int = 0) -> None:
        slot = self._next % len(self._bufs)
        if any(sl == slot for sl, _ in self._pending):
            self._events[slot].synchronize()          # ring full: wait for the oldest record only
        buf = self._bufs[slot]
        buf.fill_(float("nan"))
        if stats16 is not None:
            buf[:16].copy_(stats16, non_blocking=True)
        if grad_stats5 is not None:
            buf[16:].copy_(grad_stats5, non_blocking=True)
        self._events[slot].record()
        self._pending = [(sl, st) for sl, st in self._pending if sl != slot] + [(slot, step)]
        self._next += 1

    def pop_ready(self, wait: bool = False):
        done, keep = [], []
        for slot, step in self._pending:
            if wait:
                self._events[slot].synchronize()
            if self._events[slot].query():
                vals = self._bufs[slot].tolist()
                rec = dict(zip(ops.STAT_KEYS, vals[:16]))
                rec.update(zip(self.GRAD_KEYS, vals[16:]))
                rec["step"] = step
                done.append(rec)
            else:
                keep.append((slot, step))
        self._pending = keep
        return done
