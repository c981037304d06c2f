#!/usr/bin/env python
"""Only the "unlearn steps/s around the real UNet architecture" block of bench.py (1 GPU, or N under torchrun):

    python tools/real_unet_steps.py [--out gpurun_out/real_unet_steps_w1.json]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29561 \
        tools/real_unet_steps.py --out gpurun_out/real_unet_steps_wN.json

Full optimiser steps through the public API (UnlearnStep + GradCombiner + FusedCombineAdamW) around
tools/unet2d.py::celebahq256, per-GPU batch 16 x 3x256x256 bf16. Rank 0 prints / writes one JSON document."""
import argparse
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--transport", default="auto")
    ap.add_argument("--stand-in", action="store_true", help="the light conv stand-in instead (bench.py's `unlearn_steps`)")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from siss_b200 import _lib
    from siss_b200.scheduler import SissDDPMScheduler
    _lib.load()
    args = argparse.Namespace(channels=3, res=256, dtype="bf16", params=bench.CELEB_PARAMS, transport=a.transport)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    res = bench.unlearn_steps_with_standin(args, dev, world, rank, SissDDPMScheduler(), barrier, dist, real_arch=not a.stand_in)
    res["n_gpus"] = world
    res["peak_memory_gb"] = torch.cuda.max_memory_allocated(dev) / 1e9
    if rank == 0:
        txt = json.dumps(res, indent=1)
        print(txt)
        if a.out:
            Path(a.out).parent.mkdir(parents=True, exist_ok=True)
            Path(a.out).write_text(txt)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
