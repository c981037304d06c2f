#!/usr/bin/env python
"""One launch of every exchange schedule at full size, for an ncu NVLink-counter capture of rank 0
(tools/ncu_nvlink.sh wraps rank 0 in ncu; the other ranks run this script plainly)."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
from siss_b200 import _lib  # noqa: E402
from siss_b200.grad_combine import GradCombiner  # noqa: E402

P = int(os.environ.get("SISS_CAPTURE_P", 113_673_219))
holder = torch.nn.Parameter(torch.empty(P, device=dev))
comb = GradCombiner([holder], transport="p2p")
pe = comb.peer
g = torch.Generator(device=dev).manual_seed(rank)
for algo in pe.available():
    pe.g_x.copy_(torch.randn(comb.total, device=dev, generator=g) * 1e-3)
    pe.g_a.copy_(torch.randn(comb.total, device=dev, generator=g) * 1e-3)
    torch.cuda.synchronize(); dist.barrier()
    pe.combine(_lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0, False, comb.stats, algo=algo)
    torch.cuda.synchronize(); dist.barrier()
    if rank == 0:
        print("ran", algo, flush=True)
dist.barrier()
dist.destroy_process_group()
