"""Two-term gradient combine (K4) — the function the reference never factored out.

In the reference the logic is inline, three times (delete_celeb.py:682-767, delete_tshirt.py:624-711,
delete_sd.py:1039-1123):

    backward(weighted_loss_x, retain_graph)          ; clone every param.grad      -> g_x
    backward(weighted_loss_a)                        ; grad - g_x, += into a dict  -> accum_a
    at the sync step: accum_x = grad - accum_a ; ||accum_x||, ||accum_a|| ; scaling_factor ;
                      param.grad = accum_x - scaling_factor * accum_a ; clip_grad_norm_(1.0)

B200 design: two flat fp32 buffers ``G_x`` and ``G_a`` (P parameters each, resident in HBM for the
whole run). ``param.grad`` is a *view* into one of them; :meth:`begin_x` / :meth:`begin_a` re-point
the views so autograd itself accumulates the keep term into ``G_x`` and the forget (NegGrad) term
into ``G_a`` — the reference's per-micro-step clone / subtract / ``+=`` passes (10 P-sized passes and
~4 launches per parameter tensor) disappear. At the sync step :meth:`combine` runs K4a
(``siss_norm3``: three sums in one 8 B/param pass) and K4b (``siss_combine``: scale, subtract and the
folded ``clip_grad_norm_`` in one 12 B/param pass). The scalars stay on the device; nothing here
synchronises the host.

Data parallel (one process per GPU): samples are independent, so the only exchange is the gradient
sum. With a process group, :meth:`combine` sums ``G_x`` and ``G_a`` over the ranks into 1/N shards, runs
K4a on the shard, exchanges the three fp64 scalars the scaling and the clip need, runs K4b on the shard
and gathers the result — as fused kernels over NVLink peer memory, NVSwitch multicast or the copy
engines (siss_b200/p2p.py) or as NCCL collectives, whichever a start-up measurement finds fastest.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist

from . import ops
from ._lib import SISS_COMBINE_ERASEDIFF, SISS_COMBINE_NONE, SISS_COMBINE_SCALING_NORM

_ALIGN = 4  # parameters start on 16-byte boundaries inside the flat buffers


class GradCombiner:
    def __init__(self, params: Iterable[torch.nn.Parameter], process_group: Optional[dist.ProcessGroup] = None,
                 distributed: Optional[bool] = None, transport: str = "auto", overlap_regions: Optional[int] = None):
        """``transport`` selects how the data-parallel exchange is carried when world > 1:
        ``"p2p"`` fused peer-memory kernels over NVLink (csrc/p2p.cu); ``"nvls"`` the same through the NVSwitch's
        in-fabric reduction / replication (multimem, csrc/nvls.cu); ``"ce"`` bytes moved by the copy engines, SMs on
        local memory only (csrc/ce.cu); ``"pipe"`` / ``"pipe_nvls"`` / ``"pipe_ce"`` the pipelined schedule (G_a first,
        then reduce + combine + broadcast of G_x in one multicast kernel, clip afterwards; scaling-norm modes);
        ``"nccl"`` torch.distributed collectives around K4a/K4b (one grouped reduce-scatter);
        ``"auto"`` every schedule the node supports AND the NCCL collectives are timed on the real buffers at
        construction (max over ranks, so the choice is collective) and the fastest is kept — separately for the
        full exchange and for the exchange whose G_x shard was reduced early (siss_b200/p2p.py::PeerExchange.tune).

        ``overlap_regions`` (default: ``SISS_OVERLAP_REGIONS`` or 1 = off; peer / multicast / DMA transports): cut the
        flat buffers into that many regions and issue the reduce of ``G_a`` region by region UNDER the second backward
        pass, driven by post-accumulate hooks (see :meth:`begin_a`). Opt-in: with the stand-in UNets of this repository
        it was measured slower than the plain exchange — the kernels it hides are link-bound but still
        occupy every SM next to the backward pass."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("GradCombiner needs at least one parameter that requires grad")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32:
                raise ValueError("GradCombiner expects fp32 parameters (the reference keeps fp32 master weights; "
                                 "mixed precision is autocast, delete_celeb.py:102-108)")
            if p.device != dev:
                raise ValueError("all parameters must live on one device")
        self.device = dev
        if distributed is None:
            distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1
        self.group = process_group
        self.world = dist.get_world_size(process_group) if distributed else 1
        self.rank = dist.get_rank(process_group) if distributed else 0

        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.num_params = sum(p.numel() for p in self.params)
        # Regions: the flat buffers are cut into R equal regions whose G_a reduce is issued as soon as autograd has
        # finalised them (begin_a); rank r owns slice r of every region. R = 1 switches the overlap off.
        req = int(os.environ.get("SISS_OVERLAP_REGIONS", "1")) if overlap_regions is None else int(overlap_regions)
        self._regions_req = max(1, min(64, req)) if self.world > 1 else 1
        quantum = _ALIGN * self.world * self._regions_req
        self.total = (off + quantum - 1) // quantum * quantum  # padded so every rank's slice of every region is 16B aligned
        self.peer = None
        self.tuning = {}
        peer_names = ("p2p", "nvls", "ce", "pipe", "pipe_nvls", "pipe_ce")
        if transport not in ("auto", "nccl") + peer_names:
            raise ValueError(f"unknown transport {transport!r}")
        want_peer = transport != "nccl"
        if self.world > 1 and want_peer and dev.type == "cuda" and self.world in (2, 4, 8):
            err = None
            try:
                from .p2p import PeerExchange
                self.peer = PeerExchange(self.total, dev, process_group)
            except Exception as e:
                if transport != "auto":
                    raise
                err = e
                self.peer = None
            # The choice must be COLLECTIVE: a rank that fell back on its own would issue NCCL collectives while its
            # peers launch peer-memory kernels and wait in symmetric-memory barriers — a deadlock at the first combine().
            ok = torch.tensor([1 if self.peer is not None else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)
            if int(ok.item()) == 0:
                if transport != "auto":
                    raise RuntimeError(f"transport {transport!r} could not be set up on every rank ({err!r})")
                import warnings
                warnings.warn("siss_b200: peer-memory transport unavailable on at least one rank "
                              f"({err!r}); every rank uses NCCL collectives")
                self.peer = None
            elif transport != "auto":
                if transport not in self.peer.available():
                    raise RuntimeError(f"transport {transport!r} needs NVSwitch multicast, which this node does not provide")
                self.peer.algo = transport
                three = transport if transport in ("p2p", "nvls", "ce") else \
                    {"pipe": "p2p", "pipe_nvls": "nvls", "pipe_ce": "ce"}[transport]
                self.peer.algo3 = self.peer.algo_xpre = three
        elif self.world > 1 and transport in peer_names:
            raise RuntimeError(f"transport {transport!r} needs CUDA and 2, 4 or 8 ranks on one node")
        if self.peer is not None:
            self.g_x, self.g_a = self.peer.g_x, self.peer.g_a
        else:
            self.g_x = torch.zeros(self.total, dtype=torch.float32, device=dev)
            self.g_a = torch.zeros(self.total, dtype=torch.float32, device=dev)
        self.transport = "single" if self.world == 1 else (self.peer.algo if self.peer is not None else "nccl")
        self._nccl_full = self._nccl_xpre = self.peer is None     # which situations go through NCCL collectives
        self._views_x = [self.g_x[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self._views_a = [self.g_a[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]
        self.sums3 = torch.zeros(3, dtype=torch.float64, device=dev)
        self.stats = torch.zeros(5, dtype=torch.float32, device=dev)
        if self.world > 1:
            self.shard_len = self.total // self.world
            if self.peer is not None:
                self._shard_x, self._shard_a = self.peer.shard_x, self.peer.shard_a
            else:
                self._shard_x = torch.empty(self.shard_len, dtype=torch.float32, device=dev)
                self._shard_a = torch.empty(self.shard_len, dtype=torch.float32, device=dev)
        self._early_x = False   # G_x's reduce-scatter was already issued on the side stream (begin_a(last_micro_step=True))
        self._side = None
        self._early_done = None
        self._dirty_x = False  # G_x holds the previous combined gradient and must be cleared
        # compute back-ends: the CUDA kernels. (tests of the collective choreography on gloo/CPU
        # replace these two attributes with the oracle; the product has no other path.)
        self._norm3 = ops.norm3
        self._combine = ops.combine
        if self.world > 1 and transport == "auto" and self.peer is not None:
            self._autotune()
        # Overlap of the G_a reduce with the second backward pass (peer / multicast / DMA transports): see begin_a().
        self._early_a = False          # this optimiser step's G_a reduce is being issued region by region on the side stream
        self._armed = False
        self.regions = 1
        if self.peer is not None and not self._nccl_xpre and self._regions_req > 1 and os.environ.get("SISS_NO_OVERLAP") != "1":
            self.regions = self._regions_req
            self.peer.set_regions(self.regions)
            per = self.total // self.regions
            self._param_regions = [list(range(o // per, min((o + max(p.numel(), 1) - 1) // per, self.regions - 1) + 1))
                                   for o, p in zip(self.offsets, self.params)]
            self._region_nparams = [0] * self.regions
            for js in self._param_regions:
                for j in js:
                    self._region_nparams[j] += 1
            import weakref
            me = weakref.ref(self)                 # the hooks must not keep a discarded combiner (and its buffers) alive

            def _hook(i):
                cb = me()
                if cb is not None:
                    cb._on_grad(i)
            for i, prm in enumerate(self.params):
                prm.register_post_accumulate_grad_hook(lambda _p, i=i: _hook(i))
        self._point(self._views_x)   # start out accumulating into G_x (needed by the after_backward_* spelling)

    # ------------------------------------------------------------------------------------------
    def _nccl_exchange(self, mode: int, value: float, mn: float, inf_guard: bool, x_prereduced: bool) -> None:
        """The exchange as torch.distributed collectives around K4a / K4b: reduce-scatter G_x and G_a (ONE grouped
        NCCL launch), K4a on the 1/N shard, all-reduce of the three fp64 scalars, K4b on the shard, all-gather."""
        if x_prereduced:
            dist.reduce_scatter_tensor(self._shard_a, self.g_a, op=dist.ReduceOp.SUM, group=self.group)
        elif self.device.type == "cuda":
            # one grouped NCCL launch (ncclGroupStart/End) for both buffers instead of two serialised collectives
            with dist._coalescing_manager(group=self.group, async_ops=False):
                dist.reduce_scatter_tensor(self._shard_x, self.g_x, op=dist.ReduceOp.SUM, group=self.group)
                dist.reduce_scatter_tensor(self._shard_a, self.g_a, op=dist.ReduceOp.SUM, group=self.group)
        else:
            dist.reduce_scatter_tensor(self._shard_x, self.g_x, op=dist.ReduceOp.SUM, group=self.group)
            dist.reduce_scatter_tensor(self._shard_a, self.g_a, op=dist.ReduceOp.SUM, group=self.group)
        self._norm3(self._shard_x, self._shard_a, out=self.sums3)
        dist.all_reduce(self.sums3, op=dist.ReduceOp.SUM, group=self.group)  # the scalar-norm all-reduce
        self._combine(self._shard_x, self._shard_a, self.sums3, mode, value, mn, inf_guard, out=self._shard_x,
                      stats=self.stats)
        dist.all_gather_into_tensor(self.g_x, self._shard_x, group=self.group)

    def _autotune(self) -> None:
        """transport="auto": time every schedule of the peer / multicast kernels AND the NCCL collectives on the real
        buffers (max over ranks, so every rank takes the same decision) and keep the fastest — separately for the full
        exchange and for the exchange whose G_x shard was reduced early. Results stay in ``self.tuning``."""
        res = self.peer.tune(extra={"nccl": lambda xpre: self._nccl_exchange(SISS_COMBINE_SCALING_NORM, 500.0, 1.0,
                                                                             False, xpre)})
        self.tuning = res
        best_peer_full = res.get(self.peer.algo, float("inf"))
        best_peer_xpre = res.get(self.peer.algo_xpre + "+xpre", float("inf"))
        # NCCL must beat the fused kernels by more than the tuning margin (ties go to the kernels we control)
        self._nccl_full = res.get("nccl", float("inf")) < best_peer_full * 0.97
        self._nccl_xpre = res.get("nccl+xpre", float("inf")) < best_peer_xpre * 0.97
        self.transport = "nccl" if self._nccl_full else self.peer.algo
        self.stats.zero_()

    # ------------------------------------------------------------------------------------------
    def _point(self, views: List[torch.Tensor]) -> None:
        for p, v in zip(self.params, views):
            p.grad = v

    def begin_x(self) -> None:
        """Call before ``backward(weighted_loss_x)``: gradients now accumulate into ``G_x``."""
        if self._dirty_x:
            self.g_x.zero_()
            self._dirty_x = False
        self._point(self._views_x)

    def begin_a(self, last_micro_step: bool = False) -> None:
        """Call between the two backward passes: gradients now accumulate into ``G_a``. This is the
        whole of delete_celeb.py:693-711 (clone, subtract, accumulate).

        ``last_micro_step=True`` (data parallel only) says that ``G_x`` is now FINAL for this optimiser step —
        the keep-term backward of the last accumulation micro-step has been enqueued — so its reduce-scatter is
        issued right away on a side stream and overlaps with the second backward pass; :meth:`combine` then
        only has to move ``G_a``. Do not accumulate further micro-steps into ``G_x`` after passing it."""
        self._point(self._views_a)
        if last_micro_step and self.world > 1 and self.device.type == "cuda":
            if self._side is None:
                self._side = torch.cuda.Stream(device=self.device)
                self._early_done = torch.cuda.Event()
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.device))      # backward #1 fully enqueued before this point
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                if self.regions > 1:
                    # region layout: this rank's slice of every region, back to back (one grouped NCCL launch)
                    with dist._coalescing_manager(group=self.group, async_ops=False):
                        for v in self.peer.region_views:
                            dist.reduce_scatter_tensor(self._shard_x[v["shard_off"]:v["shard_off"] + v["slice_len"]],
                                                       self.g_x[v["start"]:v["start"] + v["slice_len"] * self.world],
                                                       op=dist.ReduceOp.SUM, group=self.group)
                else:
                    dist.reduce_scatter_tensor(self._shard_x, self.g_x, op=dist.ReduceOp.SUM, group=self.group)
                self._early_done.record(self._side)
            self._early_x = True
            if self.regions > 1:
                # ... and G_a's reduce follows it on the same side stream, region by region, UNDER backward #2: a
                # post-accumulate hook on every parameter counts down the parameters overlapping each region of the flat
                # buffer; as soon as region j (j = R-1 .. 0: autograd finishes the last-registered parameters first) is
                # final, every rank enqueues a barrier for it and reduces its slice of it (PeerExchange.early_reduce_a) —
                # all ranks at once, so both directions of every link are busy. Whatever has not fired by the time the
                # exchange is called (parameters without a gradient, the first-registered layers) is flushed there, in
                # the same order on every rank.
                self._pending = list(self._region_nparams)
                self._next_region = self.regions - 1
                self._armed = True
                self._early_a = True

    # The same protocol in the "after" spelling SURVEY.md §8b sketches for this boundary: nothing to call before the
    # first backward of a micro-step (the combiner starts out, and is left by after_backward_a / zero_grad /
    # combine, pointing at G_x); zero_grad() stands in for optimizer.zero_grad() (delete_celeb.py:773), which must not
    # be called with set_to_none=True because param.grad has to stay a view of the flat buffers.
    def after_backward_x(self, last_micro_step: bool = False) -> None:
        """After ``backward(weighted_loss_x)``: what follows accumulates into ``G_a`` (== :meth:`begin_a`)."""
        self.begin_a(last_micro_step=last_micro_step)

    def after_backward_a(self) -> None:
        """After ``backward(weighted_loss_a)``: the next micro-step's keep term accumulates into ``G_x`` again."""
        self._point(self._views_x)

    def zero_grad(self) -> None:
        """After ``optimizer.step()``: clear the combined gradient (``G_a`` was cleared by :meth:`combine`)."""
        self._dirty_x = True
        self.begin_x()

    # ------------------------------------------------------------------------------------------
    def combine(self, scaling_norm: Optional[float] = None, eta: Optional[float] = None,
                max_norm: Optional[float] = 1.0, inf_guard: bool = False, *, mode: Optional[str] = None,
                value: Optional[float] = None, loss_scale: Optional[float] = None) -> torch.Tensor:
        """Sync-step combine. Exactly one of ``scaling_norm`` (SISS / No-IS, delete_celeb.py:746) or
        ``eta`` (EraseDiff, :741-742) must be given. Leaves the result in ``param.grad`` (views of
        ``G_x``) for ``optimizer.step()`` and returns the device tensor
        ``[norm_loss_x, norm_loss_a, scaling_factor, total_norm, clip_coef]`` (the first three are the
        reference's wandb scalars, :748). ``G_a`` is cleared for the next accumulation round.

        ``loss_scale`` (``mixed_precision: fp16`` runs, delete_celeb.py:104: ``scaler.get_scale()``): the buffers hold
        gradients of the SCALED loss. As in the reference, norms, scaling factor and combination are taken on them as
        they are (so the three logged scalars are the reference's fp16 values); the clip that
        ``accelerator.clip_grad_norm_`` applies AFTER unscaling is applied here to the scaled gradient with the
        threshold ``max_norm * loss_scale``, and the result is left SCALED in ``param.grad`` — the GradScaler's own
        ``unscale_`` inside ``optimizer.step()`` (which also does the inf check / step skip) finishes the job. Do not
        call ``accelerator.clip_grad_norm_`` as well. ``total_norm`` is reported unscaled. Costs nothing: same two
        kernels, different threshold."""
        if mode is not None:                  # combine(mode="scaling_norm" | "erasediff", value=...) spelling
            if mode not in ("scaling_norm", "erasediff") or value is None or scaling_norm is not None or eta is not None:
                raise ValueError('mode must be "scaling_norm" or "erasediff", with value= and without scaling_norm= / eta=')
            scaling_norm, eta = (value, None) if mode == "scaling_norm" else (None, value)
        if (scaling_norm is None) == (eta is None):
            raise ValueError("give exactly one of scaling_norm= or eta=")
        mode = SISS_COMBINE_SCALING_NORM if eta is None else SISS_COMBINE_ERASEDIFF
        value = float(scaling_norm if eta is None else eta)
        mn = 0.0 if max_norm is None else float(max_norm)
        scale = self._check_loss_scale(loss_scale)
        mn *= scale
        if self.world == 1:
            self._norm3(self.g_x, self.g_a, out=self.sums3)
            self._combine(self.g_x, self.g_a, self.sums3, mode, value, mn, inf_guard, out=self.g_x, stats=self.stats)
        else:
            self.exchange(mode, value, mn, inf_guard)
        if scale != 1.0:
            self.stats[3:4].mul_(1.0 / scale)       # total norm of the UNSCALED gradient, what clip_grad_norm_ returns
        self.g_a.zero_()
        self._dirty_x = True
        self._early_x = False
        self._point(self._views_x)
        return self.stats

    @staticmethod
    def _check_loss_scale(loss_scale: Optional[float]) -> float:
        if loss_scale is None:
            return 1.0
        scale = float(loss_scale)
        if not (scale > 0.0 and scale != float("inf")):
            raise ValueError(f"loss_scale must be a positive finite number, got {loss_scale!r}")
        return scale

    def _on_grad(self, i: int) -> None:
        """post-accumulate-grad hook of parameter i (runs on autograd's thread while backward() blocks the caller)."""
        if not self._armed:
            return
        for j in self._param_regions[i]:
            self._pending[j] -= 1
        self._issue_ready(torch.cuda.current_stream(self.device))

    def _issue_ready(self, producer: torch.cuda.Stream, flush: bool = False) -> None:
        while self._next_region >= 0 and (flush or self._pending[self._next_region] <= 0):
            ev = torch.cuda.Event()
            ev.record(producer)                   # the accumulation that completed this region is enqueued before this point
            self._side.wait_event(ev)
            self.peer.early_reduce_a(self._next_region, self._side)
            self._next_region -= 1

    def _finish_early_a(self) -> None:
        """Called when the exchange starts: flush the regions whose hooks have not fired, publish the summed norms, then
        order the caller's stream after the side stream. From here on the first stage of the exchange is done."""
        cur = torch.cuda.current_stream(self.device)
        self._issue_ready(cur, flush=True)
        self._armed = False
        self.peer.finish_early_reduce(self._side)
        self._early_done.record(self._side)
        cur.wait_event(self._early_done)

    def exchange(self, mode: int, value: float, max_norm: float, inf_guard: bool = False) -> torch.Tensor:
        """Data parallel only — the exchange step by itself: sum ``G_x`` / ``G_a`` over the ranks, K4a, K4b, result in
        every rank's ``G_x``; through whichever transport was chosen (fused peer / multicast kernels or NCCL collectives).
        Does NOT clear ``G_a`` or touch ``param.grad`` (that is :meth:`combine`). Returns the device stats tensor."""
        if self.world == 1:
            raise RuntimeError("exchange() is a data-parallel step")
        if self._early_a:
            self._finish_early_a()
        elif self._early_x:
            torch.cuda.current_stream(self.device).wait_event(self._early_done)
        use_nccl = self._nccl_xpre if self._early_x else self._nccl_full
        if use_nccl:
            self._nccl_exchange(mode, value, max_norm, inf_guard, self._early_x)
        else:
            self.peer.combine(mode, value, max_norm, inf_guard, self.stats, x_prereduced=self._early_x,
                              reduced=self._early_a)
        self._early_a = False
        return self.stats

    def wire_bytes(self, x_prereduced: bool = False) -> Dict[str, object]:
        """Bytes that cross this GPU's NVLink ports in one exchange, per direction, for the schedule in use (P = padded
        parameter count, S = P / N; the switch fetches a multicast reduce from ALL N replicas, the requester's included)."""
        N, P4 = self.world, 4 * self.total
        S4 = P4 // max(N, 1)
        name = "nccl" if (self._nccl_xpre if x_prereduced else self._nccl_full) else \
            (self.peer.algo_xpre if x_prereduced else self.peer.algo)
        nx = 0 if x_prereduced else 1
        if name in ("p2p", "ce", "nccl"):    # NCCL: ring-equivalent count (its NVLS path is internal to the library)
            out = inb = (N - 1) * S4 * (1 + nx) + (N - 1) * S4
        elif name == "nvls":
            out, inb = P4 * (1 + nx) + S4, S4 * (1 + nx) + P4
        elif name in ("pipe", "pipe_ce"):
            out, inb = (N - 1) * S4 + P4 + S4, (N - 1) * S4 + S4 + P4
        else:                                # pipe_nvls
            out, inb = P4 + P4 + S4, S4 + S4 + P4
        return {"schedule": name, "out_bytes": int(out), "in_bytes": int(inb)}

    def reduce_to_shards(self, single_term: bool = False) -> torch.Tensor:
        """Data parallel only — the first half of the exchange: reduce-scatter the gradient buffer(s) into this rank's
        1/N shard(s) (``_shard_x`` / ``_shard_a``) and all-reduce the three global sums. Returns ``sums3`` (global).
        The sharded (ZeRO-1) optimiser step of :class:`siss_b200.optim.FusedCombineAdamW` continues from here: it
        updates only this rank's shard of the parameters and all-gathers PARAMETERS instead of gradients."""
        if self.world == 1:
            raise RuntimeError("reduce_to_shards() is a data-parallel step")
        if self._early_x:
            torch.cuda.current_stream(self.device).wait_event(self._early_done)
        else:
            dist.reduce_scatter_tensor(self._shard_x, self.g_x, op=dist.ReduceOp.SUM, group=self.group)
        if not single_term:
            dist.reduce_scatter_tensor(self._shard_a, self.g_a, op=dist.ReduceOp.SUM, group=self.group)
        self._norm3(self._shard_x, self._shard_x if single_term else self._shard_a, out=self.sums3)
        dist.all_reduce(self.sums3, op=dist.ReduceOp.SUM, group=self.group)
        self._early_x = False
        return self.sums3

    def clip_only(self, max_norm: float = 1.0, loss_scale: Optional[float] = None) -> torch.Tensor:
        """Single-term methods (naive_del / simple_neg_del: ``loss is not None``, delete_celeb.py:682-684):
        gradients are in ``G_x``; only the data-parallel sum and ``clip_grad_norm_`` apply. ``loss_scale``: as in
        :meth:`combine` (threshold ``max_norm * loss_scale``, result left scaled for the GradScaler)."""
        scale = self._check_loss_scale(loss_scale)
        mn = float(max_norm) * scale
        if self.world == 1:
            self._norm3(self.g_x, self.g_x, out=self.sums3)
            self._combine(self.g_x, self.g_x, self.sums3, SISS_COMBINE_NONE, 0.0, mn, False,
                          out=self.g_x, stats=self.stats)
        else:
            dist.reduce_scatter_tensor(self._shard_x, self.g_x, op=dist.ReduceOp.SUM, group=self.group)
            self._norm3(self._shard_x, self._shard_x, out=self.sums3)
            dist.all_reduce(self.sums3, op=dist.ReduceOp.SUM, group=self.group)
            self._combine(self._shard_x, self._shard_x, self.sums3, SISS_COMBINE_NONE, 0.0, mn,
                          False, out=self._shard_x, stats=self.stats)
            dist.all_gather_into_tensor(self.g_x, self._shard_x, group=self.group)
        if scale != 1.0:
            self.stats[3:4].mul_(1.0 / scale)
        self._dirty_x = True
        self._point(self._views_x)
        return self.stats

    # ------------------------------------------------------------------------------------------
    def stats_dict(self) -> Dict[str, float]:
        """Host copy of the last combine's scalars (this DOES synchronise; call it when logging)."""
        v = self.stats.tolist()
        return {"gradient/norm_loss_x": v[0], "gradient/norm_loss_a": v[1], "gradient/scaling_factor": v[2],
                "gradient/total_norm": v[3], "gradient/clip_coef": v[4]}
