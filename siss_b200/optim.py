"""AdamW fused into the gradient combine.

``FusedCombineAdamW`` is what ``GradCombiner.combine(...)`` followed by ``torch.optim.AdamW.step()`` and
``optimizer.zero_grad()`` are in the reference loop (delete_celeb.py:714-773), as ONE pass over flat
buffers: parameters, both moments and the two gradient buffers are each read once and written once
(``siss_combine_adamw``), the combined gradient itself never goes to HBM.

Optional, same launch: the EMA shadow update the reference runs after the optimiser step
(``ema_model.step(unet.parameters())``, delete_celeb.py:776-777, ``cfg.ema.*`` keys) and a learning rate /
EMA decay read from device memory so schedules survive CUDA-graph replay (``lr_scheduler.step()``, :770).

Parameters are moved into one flat fp32 buffer (``param.data`` become views, exactly like the gradients);
the module keeps working unchanged. Under data parallel the step is ZeRO-1 by default: the gradients are reduced
into 1/N shards (NCCL or the fused NVLink kernels of ``GradCombiner``), every rank updates its shard of parameters,
moments and EMA shadow, and the PARAMETERS are gathered (by the update kernel's own peer / multicast stores when the
fused transports are in use).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch

from . import _lib, ops
from ._lib import SISS_COMBINE_ERASEDIFF, SISS_COMBINE_NONE, SISS_COMBINE_SCALING_NORM
from .grad_combine import GradCombiner


def ema_decay_at(optimization_step: int, decay: float = 0.9999, min_decay: float = 0.0, update_after_step: int = 0,
                 use_ema_warmup: bool = False, inv_gamma: float = 1.0, power: float = 2.0 / 3.0) -> float:
    """Decay factor of diffusers' ``EMAModel.get_decay`` (diffusers 0.27.2, not vendored by the reference; keys
    ``ema_max_decay / ema_inv_gamma / ema_power`` of config/train_tshirt_mnist.yaml:94-97) for the 1-based
    optimisation step: 0 on the first step (shadow := params), then warm-up or (1+s)/(10+s), clamped."""
    step = max(0, int(optimization_step) - int(update_after_step) - 1)
    if step <= 0:
        return 0.0
    cur = 1.0 - (1.0 + step / inv_gamma) ** -power if use_ema_warmup else (1.0 + step) / (10.0 + step)
    return max(min(cur, decay), min_decay)


class FusedCombineAdamW:
    def __init__(self, combiner: GradCombiner, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, ema: Optional[dict] = None,
                 device_schedule: bool = False, shard_optimizer: Optional[bool] = None):
        """``ema``: None, or the keyword arguments of :func:`ema_decay_at` (``{}`` for its defaults) to keep a flat
        shadow copy ``ema_flat`` updated inside the optimiser kernel. ``device_schedule``: keep {lr, ema_decay}
        in a device record (``d_sched``) that the kernel reads, refreshed stream-ordered by :meth:`set_schedule`
        — needed when step() is replayed from a CUDA graph with a changing learning rate / EMA warm-up.
        ``shard_optimizer`` (data parallel; default: on for every transport; ``False`` = exchange the combined
        gradient, then every rank applies the identical full update): ZeRO-1 layout — reduce-scatter the gradients, update only this rank's 1/N shard of parameters,
        moments and EMA shadow, all-gather the PARAMETERS. Same bytes on NVLink as all-gathering the combined
        gradient, but the 40 B/param optimiser pass and the optimiser state shrink by N."""
        self.combiner = combiner
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), \
            float(weight_decay)
        dev = combiner.device
        if shard_optimizer is None:
            # every data-parallel run: the update is done once, on 1/N of the parameters (N ranks == 1 rank tested at
            # 2, 4 and 8 ranks for the NCCL, peer-memory and multicast transports, tests/test_distributed_gpu.py)
            shard_optimizer = combiner.world > 1
        self.sharded = bool(shard_optimizer) and combiner.world > 1
        # sharded + peer / multicast transport: parameters live in symmetric memory and are gathered by the fused kernel
        # (peer stores or one multicast store per vector) — unless the start-up measurement found NCCL faster
        self.fused_gather = self.sharded and combiner.peer is not None and not combiner._nccl_xpre
        self.p_flat = combiner.peer.alloc_params() if self.fused_gather else \
            torch.zeros(combiner.total, dtype=torch.float32, device=dev)
        for p, off in zip(combiner.params, combiner.offsets):
            view = self.p_flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        if self.sharded:
            lo = combiner.rank * combiner.shard_len
            self.p_shard = self.p_flat[lo:lo + combiner.shard_len]          # view: updated in place, then all-gathered
        else:
            self.p_shard = self.p_flat
        self.exp_avg = torch.zeros_like(self.p_shard)
        self.exp_avg_sq = torch.zeros_like(self.p_shard)
        self.step_count = 0
        # the step count also lives on the device and is advanced on the stream, so that a CUDA graph captured
        # around step() applies the right bias corrections on every replay
        self.d_step = torch.zeros(1, dtype=torch.int64, device=dev)
        self.ema_cfg = None if ema is None else dict(ema)
        # EMAModel.__init__: shadow = clone(params). Sharded: this rank's slice only (gathered on demand).
        self.ema_flat = self.p_shard.clone() if ema is not None else None
        # The fused two-term step of a combiner that reduces region by region (GradCombiner.regions > 1) keeps this rank's
        # state in the REGION layout: slice r of every region, back to back, instead of one contiguous range. Decided at
        # the first step (the single-term path keeps the contiguous layout); the two cannot be mixed in one run.
        self._region_layout = False
        self.cur_ema_decay = 0.0
        self.d_sched = None
        if device_schedule:
            self.d_sched = torch.zeros(2, dtype=torch.float64, device=dev)
            self._sched_host = torch.zeros(2, dtype=torch.float64).pin_memory()
            self.set_schedule(lr=self.lr, ema_decay=0.0)

    def set_schedule(self, lr: Optional[float] = None, ema_decay: Optional[float] = None) -> None:
        """New learning rate and / or EMA decay for the following step() calls. With ``device_schedule`` the pair
        is copied to the device record on the current stream (no host sync), so graph replays pick it up."""
        if lr is not None:
            self.lr = float(lr)
        if ema_decay is not None:
            self.cur_ema_decay = float(ema_decay)
        if self.d_sched is not None:
            # a fresh pinned source per update: the previous async copy may not have been consumed yet
            self._sched_host = torch.tensor([self.lr, self.cur_ema_decay], dtype=torch.float64).pin_memory()
            self.d_sched.copy_(self._sched_host, non_blocking=True)

    # EMAModel.store / copy_to / restore, used around evaluation (delete_celeb.py:380-382) — not on the hot path
    def ema_copy_to_params(self) -> None:
        self._stored = self.p_flat.clone()
        if self.sharded and self._region_layout:
            import torch.distributed as dist
            cb = self.combiner
            for v in cb.peer.region_views:
                dist.all_gather_into_tensor(self.p_flat[v["start"]:v["start"] + v["slice_len"] * cb.world],
                                            self.ema_flat[v["shard_off"]:v["shard_off"] + v["slice_len"]], group=cb.group)
        elif self.sharded:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.p_flat, self.ema_flat, group=self.combiner.group)
        else:
            self.p_flat.copy_(self.ema_flat)

    def ema_restore_params(self) -> None:
        self.p_flat.copy_(self._stored)
        self._stored = None

    def _launch(self, sums3: Optional[torch.Tensor], mode: int, value: float, max_norm: float, inf_guard: bool,
                two_term: bool, shard: bool = False) -> None:
        cb = self.combiner
        g_x, g_a, n = (cb._shard_x, cb._shard_a, cb.shard_len) if shard else (cb.g_x, cb.g_a, cb.total)
        self.step_count += 1
        if self.ema_cfg is not None:
            if self.d_sched is None:
                self.cur_ema_decay = ema_decay_at(self.step_count, **self.ema_cfg)
            elif not torch.cuda.is_current_stream_capturing():
                # eager use of the device record; around graph replays the caller refreshes it with set_schedule()
                self.set_schedule(ema_decay=ema_decay_at(self.step_count, **self.ema_cfg))
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(_lib.load().siss_counter_add(ctypes.c_void_p(self.d_step.data_ptr()), 1, stream), "siss_counter_add")
        _lib.check(_lib.load().siss_combine_adamw(
            ctypes.c_void_p(g_x.data_ptr()), ctypes.c_void_p(g_a.data_ptr() if two_term else 0), n,
            ctypes.c_void_p(0 if sums3 is None else sums3.data_ptr()), int(mode), float(value), float(max_norm),
            int(bool(inf_guard)), ctypes.c_void_p(self.p_shard.data_ptr()), ctypes.c_void_p(self.exp_avg.data_ptr()),
            ctypes.c_void_p(self.exp_avg_sq.data_ptr()), self.lr, self.betas[0], self.betas[1], self.eps,
            self.weight_decay, self.step_count, ctypes.c_void_p(self.d_step.data_ptr()),
            ctypes.c_void_p(0 if self.d_sched is None else self.d_sched.data_ptr()),
            ctypes.c_void_p(0 if self.ema_flat is None else self.ema_flat.data_ptr()), float(self.cur_ema_decay),
            0 if shard else 1, ctypes.c_void_p(0),
            ctypes.c_void_p(cb.stats.data_ptr()), stream), "siss_combine_adamw")
        ops._count(2)

    def step(self, scaling_norm: Optional[float] = None, eta: Optional[float] = None,
             max_norm: Optional[float] = 1.0, inf_guard: bool = False, single_term: bool = False) -> torch.Tensor:
        """Sync-step combine + clip + AdamW + zero_grad. Give ``scaling_norm`` (SISS / No-IS), ``eta``
        (EraseDiff) or ``single_term=True`` (naive_del / simple_neg_del). Returns the device stats tensor of
        ``GradCombiner.combine``. Gradient buffers are left cleared."""
        cb = self.combiner
        mn = 0.0 if max_norm is None else float(max_norm)
        if cb.world > 1 and self.sharded:
            # ZeRO-1: gradients -> shards, global scalars, combine + clip + AdamW (+EMA) on the shard, parameters gathered
            import torch.distributed as dist
            if not single_term and (scaling_norm is None) == (eta is None):
                raise ValueError("give exactly one of scaling_norm= or eta= (or single_term=True)")
            if self.fused_gather and not single_term:
                # peer-memory transport: two fused kernels — reduce-scatter x2 + K4a | K4b + AdamW/EMA + parameter all-gather
                if cb.regions > 1 and not self._region_layout:
                    if self.step_count != 0:
                        raise RuntimeError("FusedCombineAdamW: single-term and two-term steps cannot be mixed in one run when "
                                           "the combiner reduces region by region (set SISS_OVERLAP_REGIONS=1)")
                    if self.ema_flat is not None:
                        self.ema_flat = cb.peer.shard_slices(self.p_flat)       # shadow = params, in the region layout
                    self._region_layout = True
                self.step_count += 1
                if self.ema_cfg is not None:
                    if self.d_sched is None:
                        self.cur_ema_decay = ema_decay_at(self.step_count, **self.ema_cfg)
                    elif not torch.cuda.is_current_stream_capturing():
                        self.set_schedule(ema_decay=ema_decay_at(self.step_count, **self.ema_cfg))
                stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
                _lib.check(_lib.load().siss_counter_add(ctypes.c_void_p(self.d_step.data_ptr()), 1, stream), "siss_counter_add")
                ops._count()
                if cb._early_a:
                    cb._finish_early_a()          # G_a's first stage ran shard by shard under backward #2
                elif cb._early_x:
                    torch.cuda.current_stream(cb.device).wait_event(cb._early_done)
                mode = SISS_COMBINE_SCALING_NORM if eta is None else SISS_COMBINE_ERASEDIFF
                cb.peer.adamw_allgather(mode, float(scaling_norm if eta is None else eta), mn, inf_guard, cb.stats,
                                        self.exp_avg, self.exp_avg_sq, self.ema_flat, self.lr, self.betas, self.eps,
                                        self.weight_decay, self.step_count, self.d_step, self.d_sched,
                                        self.cur_ema_decay, x_prereduced=cb._early_x, reduced=cb._early_a)
                cb._early_x = False
                cb._early_a = False
                cb.g_x.zero_(); cb.g_a.zero_()
                cb._dirty_x = False
                cb._point(cb._views_x)
                return cb.stats
            if self._region_layout:
                raise RuntimeError("FusedCombineAdamW: single-term and two-term steps cannot be mixed in one run when the "
                                   "combiner reduces region by region (set SISS_OVERLAP_REGIONS=1)")
            sums = cb.reduce_to_shards(single_term)
            if single_term:
                self._launch(sums, SISS_COMBINE_NONE, 0.0, mn, False, two_term=False, shard=True)
            else:
                mode = SISS_COMBINE_SCALING_NORM if eta is None else SISS_COMBINE_ERASEDIFF
                self._launch(sums, mode, float(scaling_norm if eta is None else eta), mn, inf_guard, two_term=True, shard=True)
            dist.all_gather_into_tensor(self.p_flat, self.p_shard, group=cb.group)
            cb.g_x.zero_()
            if not single_term:
                cb.g_a.zero_()
            cb._dirty_x = False
            cb._point(cb._views_x)
            return cb.stats
        if cb.world > 1:
            # exchange + combine first (result in G_x on every rank), then the update on the combined gradient
            if single_term:
                cb.clip_only(mn)
            else:
                cb.combine(scaling_norm=scaling_norm, eta=eta, max_norm=max_norm, inf_guard=inf_guard)
            self._launch(None, SISS_COMBINE_NONE, 0.0, 0.0, False, two_term=False)
            cb._dirty_x = False          # G_x cleared by the kernel; G_a by combine()
            cb._point(cb._views_x)
            return cb.stats
        if single_term:
            ops.norm3(cb.g_x, cb.g_x, out=cb.sums3)
            self._launch(cb.sums3, SISS_COMBINE_NONE, 0.0, mn, False, two_term=False)
        else:
            if (scaling_norm is None) == (eta is None):
                raise ValueError("give exactly one of scaling_norm= or eta= (or single_term=True)")
            mode = SISS_COMBINE_SCALING_NORM if eta is None else SISS_COMBINE_ERASEDIFF
            ops.norm3(cb.g_x, cb.g_a, out=cb.sums3)
            self._launch(cb.sums3, mode, float(scaling_norm if eta is None else eta), mn, inf_guard, two_term=True)
        cb._dirty_x = False
        cb._point(cb._views_x)
        return cb.stats
