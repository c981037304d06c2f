"""NVLink transports for the data-parallel gradient combine: peer memory and NVSwitch multicast (NVLS).

``G_x`` and ``G_a`` live in symmetric memory (``torch.distributed._symmetric_memory``: every rank's
allocation is mapped into every process of the node, and — on an NVSwitch box — bound to one multicast
address), and the exchange + K4 run as fused kernels of libsiss_b200.so (csrc/p2p.cu, csrc/nvls.cu)
instead of five NCCL collectives around two kernels. Schedules (``algo``):

  "p2p"        barrier | siss_p2p_reduce_norm3 (reduce-scatter x2 + K4a, peer loads)   | barrier
                       | siss_p2p_combine_allgather (K4b + all-gather, peer stores)    | barrier
  "nvls"       the same three stages with multimem.ld_reduce / multimem.st (reduction and replication
               inside the switch: 8 + 4/N bytes per parameter outbound, 4 + 8/N inbound)
  "pipe"       pipelined, scaling-norm modes only (csrc/nvls.cu): G_a reduce (peer loads) | barrier |
               ONE kernel that reduces G_x in the switch, forms x - s a and replicates it in place
               (reduce traffic outbound and gather traffic inbound at the same time) | barrier |
               local clip pass
  "pipe_nvls"  as "pipe" with the G_a reduce through the switch as well
  "ce"         the three stages with the bytes moved by the COPY ENGINES (csrc/ce.cu): DMA pulls of the peers' shards
               into staging, chunked so that the rank-ordered sum / K4a of piece c overlaps the copy of piece c+1;
               K4b per piece into the own buffer + DMA pushes. DMA reaches 700-730 GB/s per direction where SM-issued
               peer or multicast traffic reaches 530-620 (profiles/r2_ce_probe_w2.json). Bit-identical to "p2p".
  "pipe_ce"    as "pipe" with the G_a reduce by DMA

Which one is fastest depends on the rank count (the in-switch forms move more bytes at N = 2 and fewer
at N = 8), so :meth:`PeerExchange.tune` measures them on the real buffers at start-up and the choice is
made collectively (max over ranks). torch is plumbing here: allocation, rendezvous (pointer exchange),
multicast binding and the stream-ordered barriers.
"""
from __future__ import annotations

import ctypes
import os
from typing import Callable, Dict, Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from ._lib import SISS_COMBINE_SCALING_NORM

THREE_STAGE = ("p2p", "nvls", "ce")
PIPELINED = ("pipe", "pipe_nvls", "pipe_ce")
ALGOS = THREE_STAGE + PIPELINED
NEEDS_MULTICAST = ("nvls",) + PIPELINED
CE_CHUNKS = max(1, min(16, int(os.environ.get("SISS_CE_CHUNKS", "4"))))
# every rank-to-rank barrier kernel gives up (device-side trap -> CUDA error) after this long instead of spinning for
# ever: a rank that died or took another code path must not leave its peers' GPUs hung
BARRIER_TIMEOUT_MS = int(os.environ.get("SISS_BARRIER_TIMEOUT_MS", "120000"))


# Preference order of the schedules when their measured times are within TUNE_MARGIN of each other: the simplest and
# most deterministic first (rank-ordered sums, fewest host calls), the copy-engine schedules — ~60 driver calls per
# exchange, the most exposed to host / driver jitter — last. A later schedule must be more than 3 % faster to be chosen.
PREFERENCE = ("p2p", "nvls", "pipe", "pipe_nvls", "ce", "pipe_ce")
TUNE_MARGIN = 0.03


def _pick(times: Dict[str, float]) -> str:
    best = None
    for name in PREFERENCE:
        if name in times and (best is None or times[name] < times[best] * (1.0 - TUNE_MARGIN)):
            best = name
    return best if best is not None else min(times, key=times.get)


class PeerExchange:
    def __init__(self, total: int, device: torch.device, group: Optional[dist.ProcessGroup] = None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world not in (2, 4, 8):
            raise RuntimeError(f"peer-memory transport supports 2, 4 or 8 ranks on one node, got {self.world}")
        if total % (4 * self.world) != 0:
            raise ValueError("flat buffer length must be a multiple of 4 * world")
        self.total = total
        self.shard_len = total // self.world
        f32, f64 = torch.float32, torch.float64
        self.g_x = symm_mem.empty(total, dtype=f32, device=device)
        self.g_a = symm_mem.empty(total, dtype=f32, device=device)
        # two slot arrays of [world][4] doubles: per-rank partial sums of the reduce kernel, and (pipelined
        # schedule) of the x-combine kernel
        self.scalars = symm_mem.empty(8 * self.world, dtype=f64, device=device)
        self.g_x.zero_(); self.g_a.zero_(); self.scalars.zero_()
        self.h_x = symm_mem.rendezvous(self.g_x, self.group)
        self.h_a = symm_mem.rendezvous(self.g_a, self.group)
        self.h_s = symm_mem.rendezvous(self.scalars, self.group)
        arr = ctypes.c_void_p * self.world
        self.ptrs_x = arr(*[int(p) for p in self.h_x.buffer_ptrs])
        self.ptrs_a = arr(*[int(p) for p in self.h_a.buffer_ptrs])
        self.ptrs_s = arr(*[int(p) for p in self.h_s.buffer_ptrs])
        self.ptrs_s2 = arr(*[int(p) + 4 * self.world * 8 for p in self.h_s.buffer_ptrs])
        assert int(self.ptrs_x[self.rank]) == self.g_x.data_ptr(), "symmetric-memory pointer table does not match"
        self.slots1 = self.scalars[:4 * self.world]
        self.slots2 = self.scalars[4 * self.world:]
        # multicast (NVLS) addresses of the two gradient buffers; 0 when the fabric / driver has none
        self.mc_x = int(getattr(self.h_x, "multicast_ptr", 0) or 0)
        self.mc_a = int(getattr(self.h_a, "multicast_ptr", 0) or 0)
        self.mc_p = 0
        self.has_multicast = bool(self.mc_x and self.mc_a)
        self.shard_x = torch.empty(self.shard_len, dtype=f32, device=device)
        self.shard_a = torch.empty(self.shard_len, dtype=f32, device=device)
        self.sums_local = torch.zeros(3, dtype=f64, device=device)
        self.ws = torch.zeros(_lib.load().siss_p2p_workspace_bytes(), dtype=torch.uint8, device=device)
        self._staging = None       # DMA landing area of the copy-engine schedules (allocated on first use)
        self.regions = 1
        self._whole = self._make_view(0, total)
        self.region_views = [self._whole]
        self.region_sums = torch.zeros(3 * 64, dtype=f64, device=device)      # per-region partial sums of this rank
        self.algo = "p2p"          # schedule used when combine() is not told otherwise (see tune())
        self.algo_xpre = "p2p"     # ... when G_x arrives already reduced (three-stage schedules only)
        self.algo3 = "p2p"         # ... best three-stage schedule of the full exchange (EraseDiff cannot be pipelined)
        self.tuning: Dict[str, float] = {}
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    # ------------------------------------------------------------------------------------------
    def _make_view(self, start: int, length: int) -> dict:
        """Pointer tables for the sub-range [start, start + length) of the flat buffers, to be passed to the kernels with
        shard_len = length / world: they address ``peer + rank * shard_len + i``, i.e. this rank's slice of the range."""
        arr = ctypes.c_void_p * self.world
        off = 4 * start
        v = {"start": start, "slice_len": length // self.world, "shard_off": start // self.world,
             "ptrs_x": arr(*[int(p) + off for p in self.h_x.buffer_ptrs]),
             "ptrs_a": arr(*[int(p) + off for p in self.h_a.buffer_ptrs]),
             "mc_x": self.mc_x + off if self.mc_x else 0, "mc_a": self.mc_a + off if self.mc_a else 0}
        if getattr(self, "h_p", None) is not None:
            v["ptrs_p"] = arr(*[int(p) + off for p in self.h_p.buffer_ptrs])
            v["mc_p"] = self.mc_p + off if self.mc_p else 0
        return v

    def set_regions(self, regions: int) -> None:
        """Cut the flat buffers into ``regions`` equal contiguous regions; rank r then owns slice r of EVERY region (its
        shard buffers hold the slices back to back). Used by GradCombiner to reduce ``G_a`` region by region under the
        second backward pass, and by the fused ZeRO-1 optimiser step, whose state then lives in this layout."""
        if regions < 1 or regions > 64 or self.total % (4 * self.world * regions) != 0:
            raise ValueError("flat buffer length must be a multiple of 4 * world * regions (1 <= regions <= 64)")
        self.regions = regions
        per = self.total // regions
        self.region_views = [self._make_view(j * per, per) for j in range(regions)] if regions > 1 else [self._whole]

    def shard_slices(self, flat: torch.Tensor) -> torch.Tensor:
        """This rank's shard of a flat buffer in the current layout (region slices back to back) — a copy."""
        return torch.cat([flat[v["start"] + self.rank * v["slice_len"]: v["start"] + (self.rank + 1) * v["slice_len"]]
                          for v in self.region_views])

    def available(self) -> Sequence[str]:
        return ALGOS if self.has_multicast else tuple(a for a in ALGOS if a not in NEEDS_MULTICAST)

    @property
    def staging(self) -> torch.Tensor:
        if self._staging is None:
            self._staging = torch.empty(2 * (self.world - 1) * self.shard_len, dtype=torch.float32, device=self.g_x.device)
        return self._staging

    def _resolve(self, algo: Optional[str], mode: int, x_prereduced: bool) -> str:
        if x_prereduced:
            algo = algo if algo in THREE_STAGE else self.algo_xpre
        elif algo is None:
            algo = self.algo
        if algo in PIPELINED and (mode != SISS_COMBINE_SCALING_NORM or x_prereduced):
            # EraseDiff's s needs <G_x, G_a>; a prereduced G_x makes the three sums available after the first kernel
            algo = self.algo_xpre if x_prereduced else self.algo3
        if algo not in ALGOS:
            raise ValueError(f"unknown exchange schedule {algo!r}")
        if algo in NEEDS_MULTICAST and not self.has_multicast:
            raise RuntimeError(f"schedule {algo!r} needs a multicast (NVLS) binding, which this node does not provide")
        return algo

    def alloc_params(self) -> torch.Tensor:
        """Flat fp32 parameter buffer in symmetric memory (same length as G_x), for the sharded optimiser step whose
        parameter all-gather is done by peer stores / multicast stores (:meth:`adamw_allgather`). Collective."""
        import torch.distributed._symmetric_memory as symm_mem
        self.p_flat = symm_mem.empty(self.total, dtype=torch.float32, device=self.g_x.device)
        self.p_flat.zero_()
        self.h_p = symm_mem.rendezvous(self.p_flat, self.group)
        self.ptrs_p = (ctypes.c_void_p * self.world)(*[int(p) for p in self.h_p.buffer_ptrs])
        assert int(self.ptrs_p[self.rank]) == self.p_flat.data_ptr(), "symmetric-memory pointer table does not match"
        self.mc_p = int(getattr(self.h_p, "multicast_ptr", 0) or 0)
        self._whole = self._make_view(0, self.total)
        self.set_regions(self.regions)
        torch.cuda.synchronize(self.g_x.device)
        dist.barrier(group=self.group)
        return self.p_flat

    # ------------------------------------------------------------------------------------------
    def _reduce(self, lib, stream, algo: str, x_mode: int, view: Optional[dict] = None, sums_ptr: Optional[int] = None) -> None:
        """First stage on the whole buffer (default) or on one region ``view`` (its sums then go to ``sums_ptr``)."""
        P = ctypes.c_void_p
        v = self._whole if view is None else view
        sx = P(self.shard_x.data_ptr() + 4 * v["shard_off"])
        sa = P(self.shard_a.data_ptr() + 4 * v["shard_off"])
        sums = P(self.sums_local.data_ptr() if sums_ptr is None else sums_ptr)
        if algo == "ce":
            _lib.check(lib.siss_ce_reduce_norm3(v["ptrs_x"], v["ptrs_a"], self.ptrs_s, self.world, self.rank, v["slice_len"],
                                                self.staging.data_ptr(), sx, sa, sums, x_mode, CE_CHUNKS,
                                                self.ws.data_ptr(), stream), "siss_ce_reduce_norm3")
        elif algo == "nvls":
            _lib.check(lib.siss_nvls_reduce_norm3(P(v["mc_x"]), P(v["mc_a"]), self.ptrs_s, self.world, self.rank,
                                                  v["slice_len"], sx, sa, sums, x_mode, self.ws.data_ptr(), stream),
                       "siss_nvls_reduce_norm3")
        else:
            _lib.check(lib.siss_p2p_reduce_norm3(v["ptrs_x"], v["ptrs_a"], self.ptrs_s, self.world, self.rank,
                                                 v["slice_len"], sx, sa, sums, x_mode, self.ws.data_ptr(), stream),
                       "siss_p2p_reduce_norm3")

    def _reduce_regions(self, lib, stream, algo: str, x_mode: int) -> None:
        """First stage region by region (the fused optimiser's layout), then the total of the per-region sums is
        published to every peer."""
        for j, v in enumerate(self.region_views):
            self._reduce(lib, stream, algo, x_mode, view=v, sums_ptr=self.region_sums.data_ptr() + 24 * j)
        self._publish(lib, stream)

    def _publish(self, lib, stream) -> None:
        _lib.check(lib.siss_publish_sums(self.region_sums.data_ptr(), self.regions, self.ptrs_s, self.world, self.rank, stream),
                   "siss_publish_sums")

    def adamw_allgather(self, mode: int, value: float, max_norm: float, inf_guard: bool, stats: torch.Tensor,
                        exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, ema_shard: Optional[torch.Tensor],
                        lr: float, betas, eps: float, weight_decay: float, step: int, d_step: torch.Tensor,
                        d_sched: Optional[torch.Tensor], ema_decay: float, x_prereduced: bool = False,
                        algo: Optional[str] = None, reduced: bool = False) -> None:
        """Two-term sync step with the ZeRO-1 update: reduce kernel as in :meth:`combine`, then K4b + AdamW/EMA on this
        rank's shard + parameter all-gather (peer stores or one multicast store per vector). New parameters land in
        every rank's ``p_flat``. Three-stage schedules only. Stream-ordered; no host synchronisation."""
        from . import ops
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = ctypes.c_void_p
        algo = algo if algo in THREE_STAGE else (self.algo_xpre if x_prereduced else self.algo3)
        if algo == "nvls" and not (self.has_multicast and self.mc_p):
            algo = "p2p"
        # "ce": DMA reduce, then the peer-store kernel for the fused update + parameter all-gather
        if not reduced:
            self.h_x.barrier(channel=0, timeout_ms=BARRIER_TIMEOUT_MS)
            if self.regions > 1:
                self._reduce_regions(lib, stream, algo, 1 if x_prereduced else 0)
            else:
                self._reduce(lib, stream, algo, 1 if x_prereduced else 0)
        self.h_x.barrier(channel=1, timeout_ms=BARRIER_TIMEOUT_MS)
        for v in self.region_views:          # one launch per region (one in all when the buffer is not cut)
            o = 4 * v["shard_off"]
            tail = (int(mode), float(value), float(max_norm), int(bool(inf_guard)),
                    P(exp_avg.data_ptr() + o), P(exp_avg_sq.data_ptr() + o), float(lr), float(betas[0]), float(betas[1]),
                    float(eps), float(weight_decay), int(step), P(d_step.data_ptr()),
                    P(0 if d_sched is None else d_sched.data_ptr()), P(0 if ema_shard is None else ema_shard.data_ptr() + o),
                    float(ema_decay), P(stats.data_ptr()), stream)
            sx, sa = P(self.shard_x.data_ptr() + o), P(self.shard_a.data_ptr() + o)
            if algo == "nvls":
                _lib.check(lib.siss_nvls_adamw_allgather(
                    sx, sa, self.scalars.data_ptr(), P(v["mc_p"]), P(self.p_flat.data_ptr() + 4 * v["start"]), self.world,
                    self.rank, v["slice_len"], *tail), "siss_nvls_adamw_allgather")
            else:
                _lib.check(lib.siss_p2p_adamw_allgather(
                    sx, sa, self.scalars.data_ptr(), v["ptrs_p"], self.world, self.rank, v["slice_len"], *tail),
                    "siss_p2p_adamw_allgather")
        self.h_x.barrier(channel=2, timeout_ms=BARRIER_TIMEOUT_MS)        # every rank's parameter shard has landed in every p_flat
        ops._count(1 + len(self.region_views))

    def early_reduce_a(self, region: int, stream: torch.cuda.Stream) -> None:
        """One step of the reduce of ``G_a`` that runs UNDER the second backward pass (GradCombiner's post-accumulate
        hooks): every rank calls this once per region, in the order regions-1 .. 0, on a side stream, as soon as the
        parameters overlapping that region have their final gradient (or at the latest when the exchange starts). A
        barrier on the side stream makes sure the region is final on EVERY rank; then every rank reduces its slice of
        it with the first-stage kernel of the tuned three-stage schedule (``x_mode`` 1: ``shard_x`` was reduced even
        earlier, on the same stream), all ranks at once — both directions of every link busy. Uses ``G_a``'s own signal
        pads, so these barriers cannot collide with the main stream's."""
        from . import ops
        lib = _lib.load()
        v = self.region_views[region]
        with torch.cuda.stream(stream):
            self.h_a.barrier(channel=region, timeout_ms=BARRIER_TIMEOUT_MS)
            self._reduce(lib, ctypes.c_void_p(stream.cuda_stream), self.algo_xpre, 1, view=v,
                         sums_ptr=self.region_sums.data_ptr() + 24 * region)
        ops._count()

    def finish_early_reduce(self, stream: torch.cuda.Stream) -> None:
        """After the last region: publish the total of the per-region sums to every peer (side stream)."""
        self._publish(_lib.load(), ctypes.c_void_p(stream.cuda_stream))

    def combine(self, mode: int, value: float, max_norm: float, inf_guard: bool, stats: torch.Tensor,
                x_prereduced: bool = False, algo: Optional[str] = None, reduced: bool = False) -> None:
        """Result lands in every rank's ``g_x``. Stream-ordered; no host synchronisation. With ``x_prereduced`` the
        reduced ``G_x`` shard is already in ``shard_x`` (early reduce-scatter overlapped with the second backward) and
        only ``G_a`` crosses NVLink in the first kernel. ``reduced``: the first stage has already run
        (:meth:`early_reduce_a`, ordered before this call on the current stream): only the scalar barrier and the
        second stage remain. ``algo`` overrides the tuned schedule."""
        from . import ops
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = ctypes.c_void_p
        algo = self._resolve(algo, int(mode), bool(x_prereduced))
        mn, ig = float(max_norm), int(bool(inf_guard))
        if not reduced:
            self.h_x.barrier(channel=0, timeout_ms=BARRIER_TIMEOUT_MS)    # every rank's G_x / G_a are complete
        if algo in PIPELINED:
            self._reduce(lib, stream, {"pipe": "p2p", "pipe_nvls": "nvls", "pipe_ce": "ce"}[algo], 2)   # phase 1: G_a, sum a^2
            self.h_x.barrier(channel=1, timeout_ms=BARRIER_TIMEOUT_MS)    # every rank's sum a^2 is in every slot array
            _lib.check(lib.siss_nvls_xcombine_bcast(P(self.mc_x), self.shard_a.data_ptr(), self.slots1.data_ptr(),
                                                    self.ptrs_s2, self.world, self.rank, self.shard_len, float(value),
                                                    ig, self.ws.data_ptr(), stream), "siss_nvls_xcombine_bcast")
            self.h_x.barrier(channel=2, timeout_ms=BARRIER_TIMEOUT_MS)    # unclipped combination complete in every G_x; second slot array filled
            _lib.check(lib.siss_scale_finalize(self.g_x.data_ptr(), self.total, self.slots1.data_ptr(),
                                               self.slots2.data_ptr(), self.world, float(value), mn, ig,
                                               stats.data_ptr(), stream), "siss_scale_finalize")
            ops._count(3)
            return
        # A G_x shard that was reduced early lives in the region layout (slice r of every region), so whenever it is used
        # both stages run once per region; otherwise on the whole shard.
        by_region = self.regions > 1 and (reduced or x_prereduced)
        if not reduced:
            if by_region:
                self._reduce_regions(lib, stream, algo, 1)
            else:
                self._reduce(lib, stream, algo, 1 if x_prereduced else 0)
        self.h_x.barrier(channel=1, timeout_ms=BARRIER_TIMEOUT_MS)        # every rank's scalar slot has been written everywhere
        views = self.region_views if by_region else [self._whole]
        for v in views:
            o = 4 * v["shard_off"]
            sx, sa = P(self.shard_x.data_ptr() + o), P(self.shard_a.data_ptr() + o)
            if algo == "ce":
                _lib.check(lib.siss_ce_combine_allgather(sx, sa, self.scalars.data_ptr(), v["ptrs_x"], self.world, self.rank,
                                                         v["slice_len"], CE_CHUNKS, int(mode), float(value), mn, ig,
                                                         stats.data_ptr(), stream), "siss_ce_combine_allgather")
                ops._count(CE_CHUNKS - 1)      # one kernel per piece on both sides
            elif algo == "nvls":
                _lib.check(lib.siss_nvls_combine_allgather(sx, sa, self.scalars.data_ptr(), P(v["mc_x"]), self.world,
                                                           self.rank, v["slice_len"], int(mode), float(value), mn, ig,
                                                           stats.data_ptr(), stream), "siss_nvls_combine_allgather")
            else:
                _lib.check(lib.siss_p2p_combine_allgather(sx, sa, self.scalars.data_ptr(), v["ptrs_x"], self.world,
                                                          self.rank, v["slice_len"], int(mode), float(value), mn, ig,
                                                          stats.data_ptr(), stream), "siss_p2p_combine_allgather")
        self.h_x.barrier(channel=2, timeout_ms=BARRIER_TIMEOUT_MS)        # every rank's shard of the result has landed in every G_x
        ops._count(1 + len(views))

    # ------------------------------------------------------------------------------------------
    def tune(self, extra: Optional[Dict[str, Callable[[bool], None]]] = None, iters: int = 7,
             candidates: Optional[Sequence[str]] = None) -> Dict[str, float]:
        """Measure every available schedule on the real buffers (CUDA events, max over ranks) and adopt the fastest
        for (a) the full exchange and (b) the exchange with G_x already reduced. ``extra`` adds competitors that are
        not schedules of this class — e.g. ``{"nccl": fn}`` with ``fn(x_prereduced)`` running the NCCL collectives —
        so that the caller can pick between transports with the same clock. Collective: every rank must call it with
        the same arguments; the decision is taken on all-reduced (MAX) timings and is therefore identical everywhere.
        Overwrites the gradient buffers (call before the first backward). Returns {name: ms}; the winners are stored in
        ``algo`` / ``algo_xpre`` (names from ``extra`` are reported but not stored)."""
        dev = self.g_x.device
        stats = torch.zeros(5, dtype=torch.float32, device=dev)
        names = [a for a in (candidates or self.available()) if a in self.available()]
        runs = []
        for a in names:
            runs.append((a, (lambda a=a: self.combine(SISS_COMBINE_SCALING_NORM, 500.0, 1.0, False, stats, algo=a))))
            if a in THREE_STAGE:
                runs.append((a + "+xpre", (lambda a=a: self.combine(SISS_COMBINE_SCALING_NORM, 500.0, 1.0, False, stats,
                                                                    x_prereduced=True, algo=a))))
        for k, fn in (extra or {}).items():
            runs.append((k, (lambda fn=fn: fn(False))))
            runs.append((k + "+xpre", (lambda fn=fn: fn(True))))
        times = torch.zeros(len(runs), dtype=torch.float64, device=dev)
        for j, (_, fn) in enumerate(runs):
            # non-trivial values: zeros would make s infinite and let the clip pass of the pipelined schedule exit early
            self.g_x.fill_(1e-3); self.g_a.fill_(1e-3); self.shard_x.fill_(1e-3)
            fn()
            torch.cuda.synchronize(dev)
            dist.barrier(group=self.group)
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
            evs[0].record()
            for i in range(iters):
                fn()
                evs[i + 1].record()
            torch.cuda.synchronize(dev)
            per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(iters))
            times[j] = per[len(per) // 2] if len(per) % 2 else 0.5 * (per[len(per) // 2 - 1] + per[len(per) // 2])   # median: immune to one stall
        dist.all_reduce(times, op=dist.ReduceOp.MAX, group=self.group)
        self.g_x.zero_(); self.g_a.zero_()
        torch.cuda.synchronize(dev)
        dist.barrier(group=self.group)
        res = {name: float(t) for (name, _), t in zip(runs, times.tolist())}
        own_full = {k: v for k, v in res.items() if k in ALGOS}
        own_xpre = {k[:-5]: v for k, v in res.items() if k.endswith("+xpre") and k[:-5] in THREE_STAGE}
        if own_full:
            self.algo = _pick(own_full)
            three = {k: v for k, v in own_full.items() if k in THREE_STAGE}
            if three:
                self.algo3 = _pick(three)
        if own_xpre:
            self.algo_xpre = _pick(own_xpre)
        self.tuning = res
        if not any(a in ("ce", "pipe_ce") for a in (self.algo, self.algo_xpre, self.algo3)):
            self._staging = None           # the DMA landing area (2 (N-1)/N P floats) is only kept if a DMA schedule won
        return res
