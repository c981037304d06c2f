"""Data-parallel plumbing for the hot path (one process per GPU, torch.distributed / NCCL).

Samples are independent in K1-K3, so the batch is sharded by rows with no collective on the data
path; the only exchange is the gradient sum inside :class:`siss_b200.grad_combine.GradCombiner`
(reduce-scatter x2, a 3-scalar all-reduce, all-gather). This module holds the pieces that keep an
N-rank run identical to a 1-rank run on the concatenated batch:

  * the loss is normalised by the GLOBAL batch (``cfg.train_batch_size``), not the shard size;
  * the Bernoulli keep-mask is drawn ONCE for the global batch with the reference's CPU call
    (losses/ddpm_deletion_loss.py:18) on every rank — same seed, same draw — and sliced.

The reference's own data parallelism is degenerate (every rank sees the same data and seed, and the
NegGrad term is never reduced) and is deliberately not reproduced.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """Initialise torch.distributed from the torchrun environment. Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, world, local_rank


def shard_bounds(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows [lo, hi) of the global batch owned by ``rank`` (contiguous, remainder to the low ranks)."""
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rows(t: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def global_keep_mask(global_batch: int, lambd: float, rank: int, world: int) -> torch.Tensor:
    """The reference's draw for the whole batch, then this rank's slice. Every rank must call it at
    the same point of its CPU RNG stream (same seed on all ranks, as the reference's set_seed does)."""
    full = torch.rand(global_batch) > lambd
    lo, hi = shard_bounds(global_batch, rank, world)
    return full[lo:hi]
