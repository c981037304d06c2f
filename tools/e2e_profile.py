#!/usr/bin/env python
"""Kernel-time breakdown of the e2e step of bench.py (torch.profiler, CUDA activities)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from torch.profiler import ProfilerActivity, profile
import bench
from siss_b200.feed import DeviceFeeder
from siss_b200.grad_combine import GradCombiner
from siss_b200.scheduler import SissDDPMScheduler
from siss_b200.step import UnlearnStep, batch_stats

dev = torch.device("cuda", 0)
B, shape, D, P = 64, (64, 3, 256, 256), 3 * 256 * 256, bench.CELEB_PARAMS
dt = torch.bfloat16
x0_h, a0_h = bench.synth_images(shape, dt, 42)
x0_p, a0_p = x0_h.pin_memory(), a0_h.pin_memory()
unet = bench.BenchUNet(P).to(dev)
comb = GradCombiner(unet.parameters())
sched = SissDDPMScheduler()
step = UnlearnStep(unet, sched, comb, loss_fn="importance_sampling_with_mixture", train_batch_size=B, lambd=0.5,
                   scaling_norm=500.0, max_norm=1.0)
host_out = torch.empty(21).pin_memory()
feeder = DeviceFeeder([shape, shape], [dt, dt], dev)
feeder.submit([x0_p, a0_p])
done = torch.cuda.Event()


def e2e_step():
    x0, a0 = feeder.next(); feeder.submit([x0_p, a0_p])
    nz = torch.randn(shape, dtype=dt, device=dev)
    ts = torch.randint(999, 1000, (B,), device=dev).long()
    out = step.micro_step(x0, a0, nz, ts)
    bs = batch_stats(out, D)
    st = step.sync_step()
    host_out[:5].copy_(st, non_blocking=True); host_out[5:].copy_(bs, non_blocking=True)
    done.record(); done.synchronize()


for _ in range(5):
    e2e_step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(10):
        e2e_step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
