"""Opt-in device-side draws for the unlearning step.

The reference draws the noise and the timesteps with torch's device generator (delete_celeb.py:581,593) and the
Bernoulli keep/forget mask on the CPU (``torch.rand(batch_size) > lambd``, losses/ddpm_deletion_loss.py:18 — a host
RNG call, a B-byte H2D copy and, in the reference, two syncs per micro-step). Under its launcher every rank uses the
same seed and therefore the same draws.

``DeviceRng`` replaces the three draws with a counter-based stream (Philox4x32-10 + Box-Muller, ``csrc/philox.cuh``;
CPU restatement ``oracle/philox.py``): each value is a pure function of ``(seed, draw, global index)``. Consequences:
no host work and no H2D for the mask; ranks that pass their global row offset draw exactly their slice of the
1-rank tensors (N ranks == 1 rank, by construction); a CUDA graph can replay the draws if ``draw`` is advanced.
This changes what a seed means, so nothing uses it unless asked to (``UnlearnStep(..., device_rng=DeviceRng(seed))``).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib, ops


class DeviceRng:
    def __init__(self, seed: int, row_offset: int = 0, device_counter: Optional[torch.device] = None):
        """``row_offset``: this rank's first global row (``parallel.shard_bounds``) — 0 on a single GPU.
        ``device_counter``: a CUDA device on which to ALSO keep the draw index (``d_draw``); the kernels then read it
        from there and :meth:`next_draw` advances it on the stream, so a CUDA graph captured around
        ``UnlearnStep.micro_step`` draws fresh noise / timesteps / masks on every replay (the host mirror ``draw``
        only counts calls made outside of replays)."""
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.row_offset = int(row_offset)
        self.draw = 0                      # advanced once per micro-step by next_draw()
        self.d_draw = None if device_counter is None else torch.zeros(1, dtype=torch.int64, device=device_counter)

    def next_draw(self) -> int:
        """The draw index for this micro-step; call advance() once all of the micro-step's draws are enqueued."""
        return self.draw

    def advance(self) -> None:
        self.draw += 1
        if self.d_draw is not None:
            _lib.check(_lib.load().siss_counter_add(ctypes.c_void_p(self.d_draw.data_ptr()), 1, ops._stream()),
                       "siss_counter_add")
            ops._count()

    def randn(self, shape: Sequence[int], dtype: torch.dtype, device, draw: Optional[int] = None,
              elem_offset: Optional[int] = None) -> torch.Tensor:
        """N(0,1) tensor for rows [row_offset, row_offset + shape[0]) of the global batch."""
        out = torch.empty(tuple(shape), dtype=dtype, device=device)
        ops._need_cuda(out)
        if out.numel() == 0:
            return out
        per_row = out.numel() // out.shape[0]
        off = self.row_offset * per_row if elem_offset is None else int(elem_offset)
        _lib.check(_lib.load().siss_randn(ops._ptr(out), out.numel(), ops._dt(out), self.seed,
                                          self.draw if draw is None else int(draw),
                                          ops._ptr(self.d_draw if draw is None else None), off, ops._stream()), "siss_randn")
        ops._count()
        return out

    def rand_aux(self, shape: Sequence[int], dtype: torch.dtype, device, draw: Optional[int] = None) -> torch.Tensor:
        """The uniform [0, 1) tensor the in-kernel EraseDiff target uses for this draw (aux domain), materialised —
        for inspection and tests. Implemented with the kernel itself: zero predictions, target written out."""
        z = torch.zeros(tuple(shape), dtype=dtype, device=device)
        if z.numel() == 0:
            return z
        per_row = z.numel() // z.shape[0]
        out = ops.dual_mse_rng_fwd_bwd(z, z, z, 0.0, 0.0, self.seed, self.draw if draw is None else int(draw),
                                       elem_offset=self.row_offset * per_row,
                                       d_draw=self.d_draw if draw is None else None, want_target=True)
        return out[4]

    def draw_rows(self, B: int, device, t_range: Optional[Tuple[int, int]] = None, lambd: Optional[float] = None,
                  draw: Optional[int] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
        """(timesteps int64 [B] uniform in [t_lo, t_hi), keep_mask uint8 [B] = uniform > lambd); either may be
        skipped by passing None for its parameter."""
        if t_range is None and lambd is None:
            raise ValueError("give t_range= and / or lambd=")
        ts = torch.empty(B, dtype=torch.int64, device=device) if t_range is not None else None
        keep = torch.empty(B, dtype=torch.uint8, device=device) if lambd is not None else None
        ops._need_cuda(ts if ts is not None else keep)
        if B:
            lo, hi = (0, 1) if t_range is None else (int(t_range[0]), int(t_range[1]))
            _lib.check(_lib.load().siss_draw_rows(ops._ptr(ts), ops._ptr(keep), B, self.seed,
                                                  self.draw if draw is None else int(draw),
                                                  ops._ptr(self.d_draw if draw is None else None), self.row_offset, lo, hi,
                                                  0.0 if lambd is None else float(lambd), ops._stream()), "siss_draw_rows")
            ops._count()
        return ts, keep
