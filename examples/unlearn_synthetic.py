#!/usr/bin/env python
"""End-to-end synthetic unlearning run on one B200, driven by a config with the reference's key names.

    python examples/unlearn_synthetic.py [--config examples/config/delete_tshirt_like.yaml]
                                          [--loss-fn erasediff] [--steps 20]

Data, model and checkpoints are synthetic / random-init (there is no network in this image): the point is
the step the reference's delete_*.py run() bodies execute — noise, add_noise, loss_fn, two backward
passes, gradient combine, clip, AdamW — on the siss_b200 fast path with sync-free logging."""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from siss_b200 import config as siss_config
from siss_b200.grad_combine import GradCombiner
from siss_b200.optim import FusedCombineAdamW
from siss_b200.scheduler import SissDDPMScheduler
from siss_b200.step import StepLog, batch_stats


class SmallUNet(torch.nn.Module):
    """Stand-in for diffusers.UNet2DModel with its call convention (x, t, return_dict=False) -> (eps,)."""

    def __init__(self, ch=1, width=32):
        super().__init__()
        self.temb = torch.nn.Embedding(1000, width)
        self.c1 = torch.nn.Conv2d(ch, width, 3, padding=1)
        self.c2 = torch.nn.Conv2d(width, width, 3, padding=1)
        self.c3 = torch.nn.Conv2d(width, ch, 3, padding=1)

    def forward(self, x, timesteps, return_dict=False, **kw):
        h = torch.nn.functional.silu(self.c1(x.float()) + self.temb(timesteps)[:, :, None, None])
        return (self.c3(torch.nn.functional.silu(self.c2(h))),)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default=str(Path(__file__).resolve().parent / "config" / "delete_tshirt_like.yaml"))
    ap.add_argument("--loss-fn", default=None, help="override deletion.loss_fn")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    cfg = siss_config.load_config(args.config)
    if args.loss_fn:
        cfg["deletion"]["loss_fn"] = args.loss_fn
        if args.loss_fn == "simple_neg_del":
            cfg["deletion"]["loss_params"] = {"superfactor": 0.03}
        elif args.loss_fn != "importance_sampling_with_mixture":
            cfg["deletion"]["loss_params"] = {}
    hp = siss_config.hot_path_params(cfg)
    dev = torch.device("cuda", 0)
    torch.manual_seed(hp["random_seed"] or 0)
    sc = cfg["scheduler"]
    sched = SissDDPMScheduler(sc["num_train_timesteps"], float(sc["beta_start"]), float(sc["beta_end"]), sc["beta_schedule"])
    unet = SmallUNet().to(dev)
    comb = GradCombiner(unet.parameters())
    step = siss_config.step_from_config(cfg, unet, sched, comb, inf_guard=True)
    opt = FusedCombineAdamW(comb, **siss_config.adamw_kwargs(cfg))
    log = StepLog()
    B, G = hp["train_batch_size"], hp["gradient_accumulation_steps"]
    shape = (B, 1, 28, 28)
    keep_set = torch.rand(512, 1, 28, 28, device=dev) * 2 - 1          # synthetic "all" dataset
    forget = (torch.rand(1, 1, 28, 28, device=dev) * 2 - 1).expand(B, -1, -1, -1).contiguous()   # the sample to unlearn
    single = hp["loss_fn"] in ("naive_del", "simple_neg_del")
    for it in range(args.steps):
        for _ in range(G):
            x0 = keep_set[torch.randint(0, 512, (B,), device=dev)]
            noise = torch.randn(shape, device=dev)
            t = torch.randint(0, sc["num_train_timesteps"], (B,), device=dev).long()
            out = step.micro_step(x0, forget, noise, t)
        stats = batch_stats(out, 784)
        step._micro = 0
        gstats = opt.step(scaling_norm=step.scaling_norm, eta=step.eta, max_norm=1.0, inf_guard=True, single_term=single)
        log.push(stats, gstats, it)
        for rec in log.pop_ready():
            print({k: round(v, 5) for k, v in rec.items() if k in ("step", "loss_x/mean", "loss_a/mean",
                                                                   "importance_weight_x/mean", "gradient/norm_loss_a",
                                                                   "gradient/scaling_factor", "gradient/clip_coef")})
    for rec in log.pop_ready(wait=True):
        print({k: round(v, 5) for k, v in rec.items() if k in ("step", "loss_x/mean", "loss_a/mean", "gradient/clip_coef")})
    print(f"done: {args.steps} optimiser steps of {hp['loss_fn']} (B={B}, G={G})")


if __name__ == "__main__":
    main()
