"""Multi-GPU, FULL SIZE (P = 113.67 M parameters, the delete_celeb UNet): every exchange schedule — fused peer-memory
kernels, NVSwitch multicast kernels (three-stage and pipelined), NCCL collectives and the auto-tuned choice — against a
1-rank evaluation  sum_r G^(r)  ->  siss_norm3  ->  siss_combine  of the same per-rank random gradients.

What must hold:
  * "p2p" and "ce" (peer loads / DMA pulls, fixed rank-order sums): output BIT-IDENTICAL to clip * (X - s A) of the rank-ordered sums, for
    the full exchange and with the G_x shard reduced beforehand (x_prereduced), SISS and EraseDiff scalars;
  * multicast / NCCL schedules (the fabric or NCCL chooses the order of the N fp32 additions): element-wise within
    2e-6 of the largest gradient entry, scalars (norms, s, total norm, clip) rtol 2e-6;
  * every schedule: all ranks end with bit-identical buffers and statistics.
The vectorised multi-CTA paths only exist at this size (shards of 14-57 M floats); the small-network test in
test_distributed_gpu.py covers the step-level plumbing.
"""
import os
import sys
from pathlib import Path

import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

pytestmark = pytest.mark.gpu

P_CELEB = 113_673_219


def _reference(allx, alla, mode, value, max_norm):
    """1-rank evaluation on this GPU: rank-ordered fp32 sums, K4a, K4b."""
    from siss_b200 import ops
    X, A = allx[0].clone(), alla[0].clone()
    for r in range(1, allx.shape[0]):
        X += allx[r]
        A += alla[r]
    sums = ops.norm3(X, A)
    out, stats = ops.combine(X, A, sums, mode, value, max_norm)
    return X, A, out, stats


def _check(name, out, stats, X, A, ref_out, ref_stats, exact, rank, world, errs):
    import torch.distributed as dist
    torch.cuda.synchronize()
    s, clip = stats[2], stats[4]
    scale = float(ref_out.abs().max())
    err = float((out - ref_out).abs().max()) / scale
    srel = float(((stats - ref_stats).abs() / ref_stats.abs().clamp_min(1e-30)).max())
    if srel > 2e-6:
        errs.append(f"{name}: statistics {stats.tolist()} vs 1-rank {ref_stats.tolist()} (rel {srel:.2e})")
    if err > 2e-6:
        errs.append(f"{name}: max |out - ref| / max |ref| = {err:.3e}")
    if exact:
        own = (X - s * A) * clip                       # mul, sub, mul in fp32: the reference's three roundings
        if not torch.equal(out, own):
            errs.append(f"{name}: not bit-identical to clip * (X - s A) of the rank-ordered sums "
                        f"(max diff {float((out - own).abs().max()):.3e})")
    # all ranks identical, bit for bit
    mine = torch.cat([out.view(torch.int32).sum(dtype=torch.int64).reshape(1), stats.view(torch.int32).to(torch.int64)])
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    if not torch.equal(lo, hi):
        errs.append(f"{name}: ranks hold different results")
    return {"max_rel_to_peak": err, "stats_rel": srel}


def _worker(rank, world, port, q, P):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as dist
    from siss_b200 import _lib, parallel
    from siss_b200.grad_combine import GradCombiner
    parallel.init_from_env("nccl")
    dev = torch.device("cuda", rank)
    errs, report = [], {}
    try:
        g = torch.Generator(device=dev).manual_seed(1000 + rank)
        gx = torch.randn(P, device=dev, generator=g) * 1e-3
        ga = torch.randn(P, device=dev, generator=g) * 1e-3 + 0.3e-3 * gx      # correlated: <X, A> != 0 for EraseDiff
        allx, alla = torch.empty(world, P, device=dev), torch.empty(world, P, device=dev)
        dist.all_gather_into_tensor(allx.view(-1), gx)
        dist.all_gather_into_tensor(alla.view(-1), ga)
        cases = {"siss": (_lib.SISS_COMBINE_SCALING_NORM, 500.0), "erasediff": (_lib.SISS_COMBINE_ERASEDIFF, 1.5)}
        refs = {k: _reference(allx, alla, m, v, 1.0) for k, (m, v) in cases.items()}
        del allx, alla
        param = torch.nn.Parameter(torch.zeros(P, device=dev))

        probe = GradCombiner([param], transport="p2p")
        schedules = list(probe.peer.available())
        report["multicast"] = bool(probe.peer.has_multicast)
        S = probe.shard_len
        for algo in schedules:
            for case, (mode, value) in cases.items():
                X, A, ref_out, ref_stats = refs[case]
                for xpre in (False, True):
                    pe = probe.peer
                    pe.g_x.zero_(); pe.g_a.zero_()
                    pe.g_x[:P].copy_(gx); pe.g_a[:P].copy_(ga)
                    if xpre:    # the rank-ordered reduced shard, as an early reduce-scatter would have left it: this
                        Xp = torch.zeros(probe.total, device=dev)       # rank's slice of every region, back to back
                        Xp[:P].copy_(X)
                        pe.shard_x.copy_(pe.shard_slices(Xp))
                        del Xp
                    pe.combine(mode, value, 1.0, False, probe.stats, x_prereduced=xpre, algo=algo)
                    used = pe._resolve(algo, mode, xpre)
                    name = f"{algo}/{case}/{'xpre' if xpre else 'full'}(ran {used})"
                    report[name] = _check(name, pe.g_x[:P], probe.stats, X, A, ref_out, ref_stats, used in ("p2p", "ce"), rank,
                                          world, errs)
        del probe
        torch.cuda.empty_cache()

        # transports through the public GradCombiner API (NCCL collectives; the auto-tuned choice), incl. the early
        # reduce-scatter of G_x that UnlearnStep issues during the second backward pass
        for transport in ("nccl", "auto", "auto/regions4"):
            # "auto/regions4": the G_a reduce issued region by region on the side stream (GradCombiner(overlap_regions=4));
            # no backward pass runs here, so every region is flushed when the exchange starts — same kernels, same layout
            comb = GradCombiner([param], transport=transport.split("/")[0], overlap_regions=4 if "/" in transport else None)
            if "/" in transport and comb.peer is not None and not comb._nccl_xpre:
                assert comb.regions == 4
            if transport == "auto":
                report["auto_tuning_ms"] = dict(comb.tuning)
                report["auto_choice"] = {"full": comb.transport, "nccl_full": comb._nccl_full, "nccl_xpre": comb._nccl_xpre,
                                         "peer_full": comb.peer.algo if comb.peer else None,
                                         "peer_xpre": comb.peer.algo_xpre if comb.peer else None}
            for case, (mode, value) in cases.items():
                X, A, ref_out, ref_stats = refs[case]
                for early in (False, True):
                    comb.begin_x()
                    comb.g_x.zero_(); comb.g_a.zero_()
                    comb.g_x[:P].copy_(gx); comb.g_a[:P].copy_(ga)
                    comb.begin_a(last_micro_step=early)
                    kw = dict(scaling_norm=value) if case == "siss" else dict(eta=value)
                    stats = comb.combine(max_norm=1.0, **kw)
                    name = f"GradCombiner[{transport}]/{case}/{'early_x' if early else 'full'}"
                    report[name] = _check(name, comb.g_x[:P], stats, X, A, ref_out, ref_stats, False, rank, world, errs)
                    assert float(comb.g_a.abs().max()) == 0.0
            del comb
            torch.cuda.empty_cache()
    except Exception as e:  # report instead of hanging the other ranks' collectives where possible
        import traceback
        errs.append(f"rank {rank}: {e!r}\n{traceback.format_exc()}")
    if rank == 0:
        q.put((errs, report))
    try:
        dist.barrier()
        dist.destroy_process_group()
    except Exception:
        pass


@pytest.mark.parametrize("world", [2, 4, 8])
def test_exchange_fullsize(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 200 + world
    P = int(os.environ.get("SISS_TEST_P", P_CELEB))
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, P)) for r in range(world)]
    for p in procs:
        p.start()
    errs, report = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    out = ROOT / "gpurun_out"
    out.mkdir(exist_ok=True)
    import json
    (out / f"exchange_fullsize_w{world}.json").write_text(json.dumps(report, indent=1))
    assert not errs, "\n".join(errs)
    for p in procs:
        assert p.exitcode == 0
