#!/usr/bin/env python
"""Launch-bound regime: optimiser-step latency of the fast path at the delete_tshirt shape
(B = 32, 1x28x28, fp32) eager vs one CUDA-graph launch, with a small conv stand-in for the UNet."""
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

from siss_b200.grad_combine import GradCombiner
from siss_b200.scheduler import SissDDPMScheduler
from siss_b200.step import UnlearnStep, batch_stats

dev = torch.device("cuda", 0)


class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = torch.nn.Conv2d(1, 8, 3, padding=1)
        self.c2 = torch.nn.Conv2d(8, 1, 3, padding=1)

    def forward(self, x, timesteps, return_dict=False, **kw):
        return (self.c2(torch.tanh(self.c1(x))),)


B, shape = 32, (32, 1, 28, 28)
net = TinyNet().to(dev)
comb = GradCombiner(net.parameters())
step = UnlearnStep(net, SissDDPMScheduler(), comb, loss_fn="importance_sampling_with_mixture", train_batch_size=B,
                   lambd=0.5, scaling_norm=5.0, max_norm=1.0, inf_guard=True)
x0, a0 = torch.rand(shape, device=dev) * 2 - 1, torch.rand(shape, device=dev) * 2 - 1
noise, t = torch.randn(shape, device=dev), torch.randint(0, 1000, (B,), device=dev)
keep = torch.rand(B, device=dev) > 0.5


def one():
    out = step.micro_step(x0, a0, noise, t, keep_mask=keep)
    s = batch_stats(out, 784)
    return s, step.sync_step()


def timeit(fn, n=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3


from siss_b200 import ops  # noqa: E402

eager_us = timeit(one)
ops.BINDING = "ctypes"
eager_ctypes_us = timeit(one)
ops.BINDING = "torch"


# host cost of the binding alone: the four hot ops back to back on resident tensors, no UNet, no autograd
def hot_ops():
    x_mix, _, _, w_x, w_a = ops.add_noise_mixture(x0, a0, noise, keep, t, step.alphas_cumprod, step.gamma, step.sigma, 0.5)
    ops.wmse_fwd_bwd(noise, x_mix, x0, a0, t, step.gamma, step.sigma, w_x, w_a, 1.0 / B, 1.0 / B)
    ops.norm3(comb.g_x, comb.g_a, out=comb.sums3)
    ops.combine(comb.g_x, comb.g_a, comb.sums3, 0, 5.0, 1.0, True, out=comb.g_x, stats=comb.stats)


def host_us(fn, n=300):
    import time
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    dt = time.perf_counter() - t0          # enqueue time only (the GPU work of these tiny shapes is shorter)
    torch.cuda.synchronize()
    return dt / n * 1e6


hot_torch_us = host_us(hot_ops)
ops.BINDING = "ctypes"
hot_ctypes_us = host_us(hot_ops)
ops.BINDING = "torch"
side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    one()
torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    one()
graph_us = timeit(g.replay)
print(json.dumps({"shape": "delete_tshirt B=32 1x28x28 fp32, conv stand-in UNet", "eager_us_per_opt_step": eager_us,
                  "eager_us_per_opt_step_ctypes_binding": eager_ctypes_us,
                  "cuda_graph_us_per_opt_step": graph_us, "speedup": eager_us / graph_us,
                  "eager_over_graph": eager_us / graph_us,
                  "four_hot_ops_host_us": {"torch_extension": hot_torch_us, "ctypes": hot_ctypes_us},
                  "note": "eager includes torch's own dispatch of the stand-in UNet forward and two backward passes"}))
