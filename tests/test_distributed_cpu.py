"""CPU, world_size 2, gloo: the N>1 path of the gradient combine (reduce-scatter x2, 3-scalar
all-reduce, all-gather) and the sharding helpers. The CUDA kernels cannot run here, so the two compute
hooks of GradCombiner are replaced by the oracle's CPU stand-ins; everything else — flat dual
buffers, 16-byte padding per rank, view re-pointing, collectives — is the product code.

Property checked: the N-rank result equals the 1-rank result on the concatenated batch."""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.c1 = torch.nn.Conv2d(1, 4, 3, padding=1)
        self.c2 = torch.nn.Conv2d(4, 1, 3, padding=1)
        self.odd = torch.nn.Parameter(torch.zeros(3))

    def forward(self, x, timesteps, return_dict=False, **kw):
        return (self.c2(torch.tanh(self.c1(x))) + self.odd.sum(),)


def _make_batch(Bg):
    g = torch.Generator().manual_seed(123)
    x0 = torch.rand(Bg, 1, 8, 8, generator=g) * 2 - 1
    a0 = torch.rand(Bg, 1, 8, 8, generator=g) * 2 - 1
    noise = torch.randn(Bg, 1, 8, 8, generator=g)
    t = torch.randint(300, 1000, (Bg,), generator=g)
    return x0, a0, noise, t


def _accumulate(net, comb, x0, a0, noise, t, keep, Bg, mode):
    """One micro-step with the ORACLE loss on this rank's shard; gradients go to G_x / G_a through the
    product's view re-pointing."""
    from oracle import siss_oracle as O
    ac = O.make_alphas_cumprod()
    loss = O.OracleDeletionLoss(*O.gamma_sigma(ac))
    all_d = {"og_latents": x0, "noisy_latents": O.add_noise(ac, x0, noise, t)}
    del_d = {"og_latents": a0, "noisy_latents": O.add_noise(ac, a0, noise, t)}
    if mode in ("siss", "shards"):
        items = loss.importance_sampling_with_mixture(net, t, noise, {}, all_d, del_d, lambd=0.5, keep_mask=keep)
    else:
        items = loss.naive_del(net, t, noise, {}, all_d, del_d)
    if mode in ("siss", "shards"):
        comb.begin_x(); (items[5].sum() / Bg).backward(retain_graph=True)
        comb.begin_a(); (items[6].sum() / Bg).backward()
    else:
        comb.begin_x(); (items[0].sum() / Bg).backward()


def _worker(rank, world, port, mode, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    torch.set_num_threads(1)
    from oracle import siss_oracle as O
    from siss_b200 import parallel
    from siss_b200.grad_combine import GradCombiner
    r, w, _ = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    Bg = 6
    net = TinyNet()
    comb = GradCombiner(net.parameters())
    assert comb.world == world and comb.total % (4 * world) == 0
    comb._norm3, comb._combine = O.norm3_cpu, O.combine_from_sums_cpu
    x0, a0, noise, t = _make_batch(Bg)
    torch.manual_seed(9)                                # same seed on every rank -> same global draw
    keep = parallel.global_keep_mask(Bg, 0.5, rank, world)
    sh = lambda v: parallel.shard_rows(v, rank, world)
    _accumulate(net, comb, sh(x0), sh(a0), sh(noise), sh(t), keep, Bg, mode)
    if mode == "shards":
        # first half of the exchange only (what the ZeRO-1 optimiser step continues from): reduced shards + global sums
        sums = comb.reduce_to_shards()
        full_x, full_a = torch.empty(comb.total), torch.empty(comb.total)
        dist.all_gather_into_tensor(full_x, comb._shard_x)
        dist.all_gather_into_tensor(full_a, comb._shard_a)
        if rank == 0:
            q.put((torch.cat([full_x, full_a]), sums.clone()))
        dist.barrier()
        dist.destroy_process_group()
        return
    stats = comb.combine(scaling_norm=5.0, max_norm=1.0) if mode == "siss" else comb.clip_only(1.0)
    flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    for g in gathered:                                   # every rank ends with the same gradient
        assert torch.equal(g, gathered[0])
    if rank == 0:
        q.put((flat.clone(), stats.clone()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["siss", "naive"])
def test_n_rank_combine_equals_single_rank(mode, world):
    from oracle import siss_oracle as O
    from siss_b200 import parallel
    from siss_b200.grad_combine import GradCombiner
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + (0 if mode == "siss" else 1) + 10 * world
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    flat2, stats2 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # single rank on the concatenated batch, same code path with world == 1
    Bg = 6
    net = TinyNet()
    comb = GradCombiner(net.parameters(), distributed=False)
    comb._norm3, comb._combine = O.norm3_cpu, O.combine_from_sums_cpu
    x0, a0, noise, t = _make_batch(Bg)
    torch.manual_seed(9)
    keep = parallel.global_keep_mask(Bg, 0.5, 0, 1)
    _accumulate(net, comb, x0, a0, noise, t, keep, Bg, mode)
    stats1 = comb.combine(scaling_norm=5.0, max_norm=1.0) if mode == "siss" else comb.clip_only(1.0)
    flat1 = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    torch.testing.assert_close(flat2, flat1, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(stats2, stats1, rtol=2e-5, atol=1e-7)

    # and the single-rank flat-buffer result equals the reference's literal loop
    ref_net = TinyNet()
    ac = O.make_alphas_cumprod()
    loss = O.OracleDeletionLoss(*O.gamma_sigma(ac))
    all_d = {"og_latents": x0, "noisy_latents": O.add_noise(ac, x0, noise, t)}
    del_d = {"og_latents": a0, "noisy_latents": O.add_noise(ac, a0, noise, t)}
    loop = O.ReferenceGradLoop(ref_net, train_batch_size=Bg)
    if mode in ("siss", "shards"):
        items = loss.importance_sampling_with_mixture(ref_net, t, noise, {}, all_d, del_d, lambd=0.5, keep_mask=keep)
        loop.micro_step(items, retain_graph=True)
        loop.sync_step(False, scaling_norm=5.0, max_norm=1.0)
    else:
        loop.micro_step(loss.naive_del(ref_net, t, noise, {}, all_d, del_d), retain_graph=False)
        loop.sync_step(True, max_norm=1.0)
    ref_flat = torch.cat([p.grad.reshape(-1) for p in ref_net.parameters()])
    torch.testing.assert_close(flat1, ref_flat, rtol=2e-4, atol=1e-6)


def test_shard_helpers():
    from siss_b200 import parallel
    for Bg, world in [(64, 8), (6, 4), (5, 2), (3, 8)]:
        spans = [parallel.shard_bounds(Bg, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == Bg
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
        torch.manual_seed(1)
        full = torch.rand(Bg) > 0.5
        parts = []
        for r in range(world):
            torch.manual_seed(1)
            parts.append(parallel.global_keep_mask(Bg, 0.5, r, world))
        assert torch.equal(torch.cat(parts), full)


def test_combine_stand_in_matches_flat_oracle():
    """The CPU stand-ins used above reproduce the oracle's literal combine (so the gloo test really
    checks reference semantics, not just self-consistency)."""
    from oracle import siss_oracle as O
    torch.manual_seed(3)
    gx, ga = torch.randn(5000) * 3e-2, torch.randn(5000) * 1e-2
    for mode, kw, val in ((0, dict(scaling_norm=5.0), 5.0), (1, dict(eta=0.05), 0.05)):
        exp, nx, na, s, tn, clip = O.combine_flat(gx, ga, max_norm=1.0, **kw)
        stats = torch.zeros(5)
        out, _ = O.combine_from_sums_cpu(gx, ga, O.norm3_cpu(gx, ga), mode, val, 1.0, stats=stats)
        torch.testing.assert_close(out, exp, rtol=2e-5, atol=1e-8)
        torch.testing.assert_close(stats, torch.stack([nx, na, s.float(), tn, clip.float()]), rtol=2e-5, atol=1e-7)


def test_two_rank_reduce_to_shards():
    """GradCombiner.reduce_to_shards (the exchange half the sharded optimiser step starts from): the gathered shards are
    the rank-summed G_x / G_a of the 1-rank run on the concatenated batch, and sums3 are their global sums."""
    from oracle import siss_oracle as O
    from siss_b200 import parallel
    from siss_b200.grad_combine import GradCombiner
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, "shards", q)) for r in range(2)]
    for p in procs:
        p.start()
    full2, sums2 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    Bg = 6
    net = TinyNet()
    comb = GradCombiner(net.parameters(), distributed=False)
    x0, a0, noise, t = _make_batch(Bg)
    torch.manual_seed(9)
    keep = parallel.global_keep_mask(Bg, 0.5, 0, 1)
    _accumulate(net, comb, x0, a0, noise, t, keep, Bg, "shards")
    n = comb.total
    pad = full2.numel() // 2 - n                         # the 2-rank buffers are padded to a multiple of 4 * world
    assert pad >= 0 and full2[n:n + pad].abs().sum() == 0
    torch.testing.assert_close(full2[:n], comb.g_x, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(full2[n + pad:2 * n + pad], comb.g_a, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(sums2, O.norm3_cpu(comb.g_x, comb.g_a), rtol=1e-5, atol=0)
    with pytest.raises(RuntimeError):
        comb.reduce_to_shards()                          # a data-parallel step: refuses on a single rank
