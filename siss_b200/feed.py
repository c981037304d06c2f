"""Host -> device batch feeding for the unlearning loop.

The reference moves each batch with a blocking ``.to(device)`` at the top of the step
(delete_celeb.py:560-564), so the PCIe copy (50 MB per step at the celeb shape) sits on the critical
path. ``DeviceFeeder`` keeps ``depth`` device-resident slots and a dedicated copy stream: while the GPU
works on batch i, batch i+1 streams in from pinned host memory. Every batch is still copied exactly
once; ordering between the copy stream and the compute stream is by CUDA events, never by host syncs.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


class DeviceFeeder:
    def __init__(self, shapes: Sequence[Tuple[int, ...]], dtypes: Sequence[torch.dtype], device: torch.device,
                 depth: int = 2):
        if depth < 2:
            raise ValueError("depth must be >= 2 to overlap copy and compute")
        self.device = device
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(device=device)
        self.slots: List[List[torch.Tensor]] = [
            [torch.empty(s, dtype=d, device=device) for s, d in zip(shapes, dtypes)] for _ in range(depth)]
        self._ready = [torch.cuda.Event() for _ in range(depth)]   # copy of slot finished
        self._free = [torch.cuda.Event() for _ in range(depth)]    # compute on slot finished
        self._submitted = 0
        self._consumed = 0
        self._in_use = -1
        cur = torch.cuda.current_stream(device)
        for ev in self._free:
            ev.record(cur)

    def submit(self, host_tensors: Sequence[torch.Tensor]) -> None:
        """Enqueue the H2D copy of one batch (pinned host tensors) on the copy stream."""
        # occupied slots = copies submitted but not yet handed out + the slot the last next() handed out, which the
        # compute stream is still reading until the following next() records its `_free` event
        occupied = (self._submitted - self._consumed) + (1 if self._in_use >= 0 else 0)
        if occupied >= self.depth:
            raise RuntimeError("DeviceFeeder: all slots are in flight (the batch returned by the last next() still "
                               "counts as in use); call next() first")
        k = self._submitted % self.depth
        for t in host_tensors:
            if not t.is_pinned():
                raise ValueError("DeviceFeeder needs pinned host tensors (use .pin_memory())")
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._free[k])          # do not overwrite a slot still being read
            for dst, src in zip(self.slots[k], host_tensors):
                dst.copy_(src, non_blocking=True)
            self._ready[k].record(self.copy_stream)
        self._submitted += 1

    def next(self) -> List[torch.Tensor]:
        """Device tensors of the oldest submitted batch; the current stream waits for its copy. The slot
        handed out by the previous ``next()`` is released for reuse at this point of the current stream."""
        if self._consumed >= self._submitted:
            raise RuntimeError("DeviceFeeder: nothing submitted")
        cur = torch.cuda.current_stream(self.device)
        if self._in_use >= 0:
            self._free[self._in_use].record(cur)
        k = self._consumed % self.depth
        cur.wait_event(self._ready[k])
        self._in_use = k
        self._consumed += 1
        return self.slots[k]
