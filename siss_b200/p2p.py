"""NVLink peer-memory transport for the data-parallel gradient combine.

``G_x`` and ``G_a`` live in symmetric memory (``torch.distributed._symmetric_memory``: every rank's
allocation is mapped into every process of the node), and the exchange + K4 run as two fused kernels of
libsiss_b200.so (csrc/p2p.cu) instead of five NCCL collectives around two kernels:

    barrier | siss_p2p_reduce_norm3 (reduce-scatter x2 + K4a, peer loads) | barrier
            | siss_p2p_combine_allgather (K4b + all-gather, peer stores)  | barrier

torch is plumbing here: allocation, rendezvous (pointer exchange) and the stream-ordered barriers.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


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
        self.scalars = symm_mem.empty(4 * self.world, dtype=f64, device=device)
        self.g_x.zero_(); self.g_a.zero_(); self.scalars.zero_()
        self.h_x = symm_mem.rendezvous(self.g_x, self.group)
        self.h_a = symm_mem.rendezvous(self.g_a, self.group)
        self.h_s = symm_mem.rendezvous(self.scalars, self.group)
        arr = ctypes.c_void_p * self.world
        self.ptrs_x = arr(*[int(p) for p in self.h_x.buffer_ptrs])
        self.ptrs_a = arr(*[int(p) for p in self.h_a.buffer_ptrs])
        self.ptrs_s = arr(*[int(p) for p in self.h_s.buffer_ptrs])
        assert int(self.ptrs_x[self.rank]) == self.g_x.data_ptr(), "symmetric-memory pointer table does not match"
        self.shard_x = torch.empty(self.shard_len, dtype=f32, device=device)
        self.shard_a = torch.empty(self.shard_len, dtype=f32, device=device)
        self.sums_local = torch.zeros(3, dtype=f64, device=device)
        self.ws = torch.zeros(_lib.load().siss_p2p_workspace_bytes(), dtype=torch.uint8, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=self.group)

    def alloc_params(self) -> torch.Tensor:
        """Flat fp32 parameter buffer in symmetric memory (same length as G_x), for the sharded optimiser step whose
        parameter all-gather is done by peer stores (:meth:`adamw_allgather`). Collective: call on every rank."""
        import torch.distributed._symmetric_memory as symm_mem
        self.p_flat = symm_mem.empty(self.total, dtype=torch.float32, device=self.g_x.device)
        self.p_flat.zero_()
        self.h_p = symm_mem.rendezvous(self.p_flat, self.group)
        self.ptrs_p = (ctypes.c_void_p * self.world)(*[int(p) for p in self.h_p.buffer_ptrs])
        assert int(self.ptrs_p[self.rank]) == self.p_flat.data_ptr(), "symmetric-memory pointer table does not match"
        torch.cuda.synchronize(self.g_x.device)
        dist.barrier(group=self.group)
        return self.p_flat

    def adamw_allgather(self, mode: int, value: float, max_norm: float, inf_guard: bool, stats: torch.Tensor,
                        exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, ema_shard: Optional[torch.Tensor],
                        lr: float, betas, eps: float, weight_decay: float, step: int, d_step: torch.Tensor,
                        d_sched: Optional[torch.Tensor], ema_decay: float, x_prereduced: bool = False) -> None:
        """Two-term sync step with the ZeRO-1 update: reduce kernel as in :meth:`combine`, then
        ``siss_p2p_adamw_allgather`` (K4b + AdamW/EMA on this rank's shard + parameter all-gather by peer stores).
        New parameters land in every rank's ``p_flat``. Stream-ordered; no host synchronisation."""
        from . import ops
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = ctypes.c_void_p
        self.h_x.barrier(channel=0)
        _lib.check(lib.siss_p2p_reduce_norm3(self.ptrs_x, self.ptrs_a, self.ptrs_s, self.world, self.rank,
                                             self.shard_len, self.shard_x.data_ptr(), self.shard_a.data_ptr(),
                                             self.sums_local.data_ptr(), int(bool(x_prereduced)), self.ws.data_ptr(),
                                             stream), "siss_p2p_reduce_norm3")
        self.h_x.barrier(channel=1)
        _lib.check(lib.siss_p2p_adamw_allgather(
            self.shard_x.data_ptr(), self.shard_a.data_ptr(), self.scalars.data_ptr(), self.ptrs_p, self.world,
            self.rank, self.shard_len, int(mode), float(value), float(max_norm), int(bool(inf_guard)),
            P(exp_avg.data_ptr()), P(exp_avg_sq.data_ptr()), float(lr), float(betas[0]), float(betas[1]), float(eps),
            float(weight_decay), int(step), P(d_step.data_ptr()), P(0 if d_sched is None else d_sched.data_ptr()),
            P(0 if ema_shard is None else ema_shard.data_ptr()), float(ema_decay), P(stats.data_ptr()), stream),
            "siss_p2p_adamw_allgather")
        self.h_x.barrier(channel=2)        # every rank's parameter shard has landed in every p_flat
        ops._count(2)

    def combine(self, mode: int, value: float, max_norm: float, inf_guard: bool, stats: torch.Tensor,
                x_prereduced: bool = False) -> None:
        """Result lands in every rank's ``g_x``. Stream-ordered; no host synchronisation. With
        ``x_prereduced`` the reduced ``G_x`` shard is already in ``shard_x`` (early reduce-scatter overlapped
        with the second backward) and only ``G_a`` crosses NVLink in the first kernel."""
        from . import ops
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        self.h_x.barrier(channel=0)        # every rank's G_x / G_a are complete
        _lib.check(lib.siss_p2p_reduce_norm3(self.ptrs_x, self.ptrs_a, self.ptrs_s, self.world, self.rank,
                                             self.shard_len, self.shard_x.data_ptr(), self.shard_a.data_ptr(),
                                             self.sums_local.data_ptr(), int(bool(x_prereduced)), self.ws.data_ptr(),
                                             stream),
                   "siss_p2p_reduce_norm3")
        self.h_x.barrier(channel=1)        # every rank's scalar slot has been written everywhere
        _lib.check(lib.siss_p2p_combine_allgather(self.shard_x.data_ptr(), self.shard_a.data_ptr(),
                                                  self.scalars.data_ptr(), self.ptrs_x, self.world, self.rank,
                                                  self.shard_len, int(mode), float(value), float(max_norm),
                                                  int(bool(inf_guard)), stats.data_ptr(), stream),
                   "siss_p2p_combine_allgather")
        self.h_x.barrier(channel=2)        # every rank's shard of the result has landed in every G_x
        ops._count(2)
