"""CPU: host-side logic of the exchange transports — schedule preference / margin rule, NVLink byte accounting per
schedule, schedule resolution (which schedules may be pipelined), and that both bench arms print one config."""
import types

import pytest
import torch


def test_pick_prefers_simple_schedules_within_margin():
    from siss_b200.p2p import PREFERENCE, TUNE_MARGIN, _pick
    assert PREFERENCE[0] == "p2p" and 0 < TUNE_MARGIN < 0.1
    # measured tables of round 2 (ms): N = 4 — everything within 3 %: the rank-ordered peer kernels stay
    assert _pick({"p2p": 1.93, "nvls": 2.13, "ce": 2.08, "pipe": 1.98, "pipe_nvls": 2.0, "pipe_ce": 1.90}) == "p2p"
    # N = 8 — the pipelined multicast schedule wins by 14 %
    assert _pick({"p2p": 2.28, "nvls": 2.09, "ce": 2.7, "pipe": 2.05, "pipe_nvls": 1.95, "pipe_ce": 2.4}) == "pipe_nvls"
    # N = 2 — DMA wins by 5.7 %
    assert _pick({"p2p": 1.279, "ce": 1.206, "nvls": 2.2, "pipe": 1.9, "pipe_nvls": 2.2, "pipe_ce": 1.85}) == "ce"
    # a later schedule that is only 2 % faster does not displace an earlier one
    assert _pick({"p2p": 1.00, "ce": 0.98}) == "p2p"
    assert _pick({"nvls": 1.0}) == "nvls"


def _fake_combiner(world, total, algo, algo_xpre="p2p", nccl_full=False, nccl_xpre=False):
    from siss_b200.grad_combine import GradCombiner
    cb = GradCombiner.__new__(GradCombiner)
    cb.world, cb.total = world, total
    cb._nccl_full, cb._nccl_xpre = nccl_full, nccl_xpre
    cb.peer = types.SimpleNamespace(algo=algo, algo_xpre=algo_xpre)
    return cb


@pytest.mark.parametrize("world", [2, 4, 8])
def test_wire_bytes_per_schedule(world):
    P = 113_673_224 // (4 * world) * (4 * world)
    P4, S4 = 4 * P, 4 * P // world
    wb = lambda **kw: _fake_combiner(world, P, **kw).wire_bytes()
    p2p = wb(algo="p2p")
    assert p2p["out_bytes"] == p2p["in_bytes"] == 3 * (world - 1) * S4          # 12 (N-1)/N bytes per parameter
    assert wb(algo="ce") == dict(p2p, schedule="ce")
    nv = wb(algo="nvls")
    assert nv["out_bytes"] == 2 * P4 + S4 and nv["in_bytes"] == 2 * S4 + P4      # 8 + 4/N out, 4 + 8/N in
    pn = wb(algo="pipe_nvls")
    assert pn["out_bytes"] == 2 * P4 + S4 and pn["in_bytes"] == 2 * S4 + P4
    pp = wb(algo="pipe")
    assert pp["out_bytes"] == pp["in_bytes"] == (world - 1) * S4 + P4 + S4      # G_a by peer loads, then 4 + 4/N each way
    # the pipelined schedule needs fewer bytes per direction than the three-stage peer kernels from N = 4 on
    assert (pp["out_bytes"] < p2p["out_bytes"]) == (world >= 4)
    # pre-reduced G_x: only G_a and the result cross the links
    xp = _fake_combiner(world, P, algo="pipe_nvls", algo_xpre="p2p").wire_bytes(x_prereduced=True)
    assert xp["schedule"] == "p2p" and xp["out_bytes"] == 2 * (world - 1) * S4
    assert _fake_combiner(world, P, algo="p2p", nccl_full=True).wire_bytes()["schedule"] == "nccl"


def test_schedule_resolution_rules():
    """EraseDiff (s needs <G_x, G_a>) and a pre-reduced G_x cannot be pipelined; multicast schedules need the binding."""
    from siss_b200 import _lib
    from siss_b200.p2p import PeerExchange
    pe = PeerExchange.__new__(PeerExchange)
    pe.algo, pe.algo_xpre, pe.algo3, pe.has_multicast = "pipe_nvls", "nvls", "p2p", True
    SN, ED = _lib.SISS_COMBINE_SCALING_NORM, _lib.SISS_COMBINE_ERASEDIFF
    assert pe._resolve(None, SN, False) == "pipe_nvls"
    assert pe._resolve(None, ED, False) == "p2p"            # best three-stage schedule of the full exchange
    assert pe._resolve(None, SN, True) == "nvls"            # best three-stage schedule with G_x pre-reduced
    assert pe._resolve("pipe", SN, True) == "nvls"
    assert pe._resolve("ce", ED, True) == "ce"              # an explicit three-stage request is honoured
    pe.has_multicast = False
    assert set(pe.available()) == {"p2p", "ce"}
    with pytest.raises(RuntimeError):
        pe._resolve("nvls", SN, False)
    with pytest.raises(ValueError):
        pe._resolve("ring", SN, False)


def test_both_bench_arms_print_one_config():
    import argparse
    import bench
    args = argparse.Namespace(batch=64, channels=3, res=256, dtype="bf16", params=bench.CELEB_PARAMS)
    for n in (1, 2, 8):
        c = bench.workload_config(args, n)
        assert "transport" not in c and c["parallelism"] == f"dp{n}" and c["global_batch"] == 64 * n
    w = bench.celeb_workload(args)
    assert w["B"] == 64 and w["chw"] == (3, 256, 256) and w["dt"] == torch.bfloat16
    assert bench.TSHIRT["B"] == 32 and bench.SD["cond"] == (77, 768)


# ---- mixed_precision: fp16 (GradScaler) -------------------------------------------------------------------------------
@pytest.mark.parametrize("kw", [dict(scaling_norm=5.0), dict(scaling_norm=500.0), dict(eta=0.05)])
def test_oracle_fp16_sequence_is_the_unscaled_run_with_scaling_norm_divided_by_the_scale(kw):
    """What the reference does under a GradScaler (oracle.combine_flat(loss_scale=), delete_celeb.py:725-767): norms and
    scaling factor are taken on SCALED gradients, so for a power-of-two scale S the result is, bit for bit, the unscaled
    run with scaling_norm / S (EraseDiff's factor is scale-invariant) — the scale dependence INTEGRATION.md §2 describes."""
    import torch
    from oracle import siss_oracle as O
    torch.manual_seed(11)
    gx, ga = torch.randn(4096) * 3e-2, torch.randn(4096) * 1e-2 + 0.1 * torch.randn(4096) * 3e-2
    S = 1024.0
    got = O.combine_flat(gx * S, ga * S, max_norm=1.0, loss_scale=S, **kw)
    kw_ref = dict(scaling_norm=kw["scaling_norm"] / S) if "scaling_norm" in kw else kw
    exp = O.combine_flat(gx, ga, max_norm=1.0, **kw_ref)
    assert torch.equal(got[0], exp[0])
    assert float(got[4]) == float(exp[4]) and float(got[5]) == float(exp[5])          # total norm, clip coefficient
    assert float(got[1]) == float(exp[1]) * S and float(got[2]) == float(exp[2]) * S  # the logged norms are the scaled ones


@pytest.mark.parametrize("kw,mode,val", [(dict(scaling_norm=5.0), 0, 5.0), (dict(eta=0.05), 1, 0.05)])
def test_grad_combiner_loss_scale_on_cpu_stand_ins(kw, mode, val):
    """GradCombiner.combine(loss_scale=S) (host logic; K4a / K4b replaced by the CPU stand-ins): the result, unscaled as the
    GradScaler does inside optimizer.step(), is the oracle's fp16 sequence; total_norm is reported unscaled; clip_only too."""
    import torch
    from oracle import siss_oracle as O
    from siss_b200.grad_combine import GradCombiner
    torch.manual_seed(5)
    prm = torch.nn.Parameter(torch.zeros(3000))
    comb = GradCombiner([prm], distributed=False)
    comb._norm3, comb._combine = O.norm3_cpu, O.combine_from_sums_cpu
    gx, ga = torch.randn(3000) * 4e-2, torch.randn(3000) * 1e-2
    S = 4096.0
    comb.g_x[:3000].copy_(gx * S); comb.g_a[:3000].copy_(ga * S)
    stats = comb.combine(max_norm=1.0, loss_scale=S, **kw).clone()
    exp, nx, na, s, tn, clip = O.combine_flat(gx * S, ga * S, max_norm=1.0, loss_scale=S, **kw)
    torch.testing.assert_close(prm.grad * (1.0 / S), exp, rtol=3e-5, atol=1e-9)
    torch.testing.assert_close(stats, torch.stack([nx, na, s.float(), tn, clip.float()]), rtol=3e-5, atol=1e-7)
    assert float(stats[4]) < 1.0, "the case must exercise the clip"
    # single-term
    comb.begin_x()
    comb.g_x[:3000].copy_(gx * S)
    stats = comb.clip_only(1.0, loss_scale=S).clone()
    tn = torch.norm(gx)
    torch.testing.assert_close(prm.grad * (1.0 / S), gx * torch.clamp(1.0 / (tn + 1e-6), max=1.0), rtol=3e-5, atol=1e-9)
    torch.testing.assert_close(stats[3], tn, rtol=3e-5, atol=0)
    with pytest.raises(ValueError):
        comb.combine(scaling_norm=5.0, loss_scale=0.0)
