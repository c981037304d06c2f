"""CUDA-graph capture of one optimiser step.

The shapes the reference actually trains at (delete_tshirt: B = 32..64 of 1x28x28; delete_sd: B = 1 of 4x64x64 latents)
are launch-bound: one optimiser step is ~10 kernels of a few microseconds each plus the UNet, and the host's enqueue time
exceeds the device time. Nothing in ``UnlearnStep.micro_step`` / ``sync_step`` / ``FusedCombineAdamW.step`` / ``batch_stats``
synchronises the host or reads pageable memory when the keep-mask is a device tensor or comes from the device RNG, so a
whole step can be captured once and replayed with ONE launch (tests/test_cuda_graph_gpu.py; bench.py's e2e
``graph_variant``: 0.27 vs 0.52 ms per step at the tshirt shape).

``CapturedStep`` packages the ritual: warm-up on a side stream (lazy initialisation, workspaces, autograd's first-call
allocations must not happen inside the capture), capture, replay. The callable must read its inputs from STATIC tensors
(fill them with ``copy_`` before ``replay()``; a ``DeviceFeeder`` slot is static) and must draw any randomness from a
``DeviceRng(device_counter=...)`` so that replays see fresh draws.
"""
from __future__ import annotations

from typing import Any, Callable

import torch


class CapturedStep:
    def __init__(self, fn: Callable[[], Any], warmup: int = 1):
        """Run ``fn`` ``warmup`` times eagerly on a side stream, then capture one call of it. ``self.outputs`` holds
        whatever ``fn`` returned during capture: static tensors that every ``replay()`` overwrites. Note that the
        warm-up calls are REAL calls (an optimiser step captured this way has been taken ``warmup`` times already)."""
        if warmup < 1:
            raise ValueError("at least one eager warm-up call is needed before capture")
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()

    def replay(self) -> Any:
        """One launch for the whole captured step, on the current stream. Returns the static output tensors."""
        self.graph.replay()
        return self.outputs
