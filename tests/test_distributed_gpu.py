"""GPU, world_size 2, NCCL: UnlearnStep + GradCombiner with the real kernels on two B200s must equal
the single-GPU run on the concatenated batch. Skipped when fewer than 2 GPUs."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu


class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.c1 = torch.nn.Conv2d(1, 4, 3, padding=1)
        self.c2 = torch.nn.Conv2d(4, 1, 3, padding=1)
        self.odd = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x, timesteps, return_dict=False, **kw):
        return (self.c2(torch.tanh(self.c1(x))) + self.odd.sum(),)


def _batch(Bg):
    g = torch.Generator().manual_seed(321)
    return (torch.rand(Bg, 1, 16, 16, generator=g) * 2 - 1, torch.rand(Bg, 1, 16, 16, generator=g) * 2 - 1,
            torch.randn(Bg, 1, 16, 16, generator=g), torch.randint(300, 1000, (Bg,), generator=g))


def _run_opt(rank, world, Bg, transport="auto", shard=None, regions=None):
    """Three optimiser steps with the fused combine+AdamW (+EMA) under data parallel; returns the final parameters
    followed by the EMA shadow parameters. ``shard``: ZeRO-1 optimiser sharding (None = the transport's default)."""
    from siss_b200 import parallel
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.optim import FusedCombineAdamW
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", rank if world > 1 else 0)
    net = TinyNet().to(dev)
    comb = GradCombiner(net.parameters(), transport=transport, overlap_regions=regions)
    opt = FusedCombineAdamW(comb, lr=3e-3, betas=(0.95, 0.999), weight_decay=1e-2, ema=dict(decay=0.9),
                            shard_optimizer=shard)
    if world > 1:
        assert opt.sharded == (True if shard is None else shard)
        # parameter all-gather inside the fused kernel (peer stores / multicast stores) whenever a peer transport is in use
        assert opt.fused_gather == (opt.sharded and comb.peer is not None and not comb._nccl_xpre)
        if opt.sharded:
            assert opt.exp_avg.numel() == comb.total // world and opt.ema_flat.numel() == comb.total // world
    step = UnlearnStep(net, SissDDPMScheduler(), comb, loss_fn="importance_sampling_with_mixture",
                       train_batch_size=Bg, lambd=0.5, scaling_norm=5.0, max_norm=1.0)
    torch.manual_seed(23)
    for it in range(3):
        x0, a0, noise, t = _batch(Bg)
        keep = parallel.global_keep_mask(Bg, 0.5, rank, world)
        sh = lambda v: parallel.shard_rows(v, rank, world).to(dev)
        step.micro_step(sh(x0 + 0.02 * it), sh(a0), sh(noise), sh(t), keep_mask=keep)
        step._micro = 0
        opt.step(scaling_norm=5.0, max_norm=1.0)
        assert comb.g_x.abs().max().item() == 0.0 and comb.g_a.abs().max().item() == 0.0
    torch.cuda.synchronize()
    final = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    opt.ema_copy_to_params()                      # sharded: all-gathers the shadow shards
    shadow = torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu()
    opt.ema_restore_params()
    assert torch.equal(torch.cat([p.detach().reshape(-1) for p in net.parameters()]).cpu(), final)
    return torch.cat([final, shadow])


def _run(rank, world, Bg, G, transport="auto", regions=None):
    from siss_b200 import parallel
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    dev = torch.device("cuda", rank if world > 1 else 0)
    net = TinyNet().to(dev)
    comb = GradCombiner(net.parameters(), transport=transport, overlap_regions=regions)
    if world > 1 and transport != "auto":
        assert comb.transport == transport
    step = UnlearnStep(net, SissDDPMScheduler(), comb, loss_fn="importance_sampling_with_mixture",
                       train_batch_size=Bg, gradient_accumulation_steps=G, lambd=0.5, scaling_norm=5.0, max_norm=1.0)
    torch.manual_seed(17)
    for k in range(G):
        x0, a0, noise, t = _batch(Bg)
        keep = parallel.global_keep_mask(Bg, 0.5, rank, world)
        sh = lambda v: parallel.shard_rows(v, rank, world).to(dev)
        step.micro_step(sh(x0 + 0.01 * k), sh(a0), sh(noise), sh(t), keep_mask=keep)
    stats = step.sync_step()
    torch.cuda.synchronize()
    return torch.cat([p.grad.reshape(-1) for p in net.parameters()]).cpu(), stats.cpu()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from siss_b200 import parallel
    parallel.init_from_env("nccl")
    out = {}
    from siss_b200.grad_combine import GradCombiner
    probe = GradCombiner(TinyNet().to(torch.device("cuda", rank)).parameters(), transport="p2p")
    multicast = probe.peer.has_multicast            # NVSwitch multicast bound: the in-switch schedules exist too
    del probe
    for transport in ("nccl", "p2p", "ce", "auto") + (("nvls", "pipe", "pipe_nvls", "pipe_ce") if multicast else ()):
        out[transport] = _run(rank, world, 8, 2, transport)
        out[transport + "/fused_adamw"] = _run_opt(rank, world, 8, transport)
        if transport in ("nccl", "p2p"):
            out[transport + "/fused_adamw_replicated"] = _run_opt(rank, world, 8, transport, shard=False)
        if transport in ("p2p", "ce") or (transport == "nvls"):
            # G_a reduced region by region under the second backward pass (post-accumulate hooks), region layout of the
            # ZeRO-1 state incl. the EMA shadow
            out[transport + "/regions2"] = _run(rank, world, 8, 2, transport, regions=2)
            out[transport + "/regions2/fused_adamw"] = _run_opt(rank, world, 8, transport, regions=2)
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_gpu_step_equals_single_gpu(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 200 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    flat1, stats1 = _run(0, 1, 8, 2)
    params1 = _run_opt(0, 1, 8)
    for transport, res in out.items():                  # NCCL collectives and fused NVLink peer-memory kernels
        if "/fused_adamw" in transport:                 # replicated and ZeRO-1 sharded layouts, parameters + EMA shadow
            torch.testing.assert_close(res, params1, rtol=3e-5, atol=3e-6, msg=lambda m: f"{transport}: {m}")
            continue
        flat2, stats2 = res
        torch.testing.assert_close(flat2, flat1, rtol=2e-4, atol=2e-6, msg=lambda m: f"{transport}: {m}")
        torch.testing.assert_close(stats2, stats1, rtol=2e-4, atol=1e-7, msg=lambda m: f"{transport}: {m}")
