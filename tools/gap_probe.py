#!/usr/bin/env python
"""Why is the un-instrumented resident step at N > 1 ~0.1 ms longer than the sum of its per-kernel brackets?
Times, on N ranks (torchrun): (a) exchange only, (b) row kernels only, (c) row kernels + exchange (the bench step),
(d) the same with an event record between the row kernels and the exchange, (e) exchange + an unrelated filler kernel."""
import json
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
from siss_b200 import _lib, ops  # noqa: E402
from siss_b200.grad_combine import GradCombiner  # noqa: E402
from siss_b200.scheduler import SissDDPMScheduler  # noqa: E402

P = 113_673_219
holder = torch.nn.Parameter(torch.empty(P, device=dev))
comb = GradCombiner([holder], transport=os.environ.get("SISS_GAP_TRANSPORT", "auto"))
g = torch.Generator(device=dev).manual_seed(rank)
comb.g_x.copy_(torch.randn(comb.total, device=dev, generator=g) * 1e-3)
comb.g_a.copy_(torch.randn(comb.total, device=dev, generator=g) * 1e-3)
B, shape = 64, (64, 3, 256, 256)
sched = SissDDPMScheduler()
ac = sched.alphas_cumprod.to(dev)
gamma, sigma = sched.gamma_sigma(dev)
x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).bfloat16()
nz = torch.randn(shape, device=dev, generator=g).bfloat16()
pred = torch.randn(shape, device=dev, generator=g)
t = torch.full((B,), 999, device=dev, dtype=torch.long)
keep = (torch.rand(B, device=dev, generator=g) > 0.5).to(torch.uint8)
filler = torch.empty(120 << 20, dtype=torch.uint8, device=dev)
SN = _lib.SISS_COMBINE_SCALING_NORM


def rows():
    x_mix, _, _, w_x, w_a = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
    ops.wmse_fwd_bwd(pred, x_mix, x0, a0, t, gamma, sigma, w_x, w_a, 1 / 64, 1 / 64)


def exch():
    comb.exchange(SN, 500.0, 1.0, False)


ev = torch.cuda.Event(enable_timing=True)
variants = {
    "exchange_only": exch,
    "rows_only": rows,
    "rows+exchange": lambda: (rows(), exch()),
    "rows+event+exchange": lambda: (rows(), ev.record(), exch()),
    "filler+exchange": lambda: (filler.zero_(), exch()),
    "exchange+sync_every_step": lambda: (rows(), exch(), torch.cuda.synchronize()),
}
res = {"world": world, "schedule": comb.transport}
for name, fn in variants.items():
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(40):
        fn()
    e.record()
    torch.cuda.synchronize()
    tt = torch.tensor([s.elapsed_time(e) / 40], device=dev, dtype=torch.float64)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    res[name] = round(float(tt.item()), 4)
if rank == 0:
    print(json.dumps(res))
dist.barrier()
dist.destroy_process_group()
