"""GPU parity for the opt-in device RNG — through the C ABI.

There is no reference output to compare with (the counter-based stream is a new seed semantic), so the chain is:
  oracle/philox.py  == Random123 known answers                  (tests/test_philox_cpu.py, CPU)
  siss_randn / siss_draw_rows == oracle/philox.py               (here: integers exact, normals to intrinsic accuracy)
  siss_add_noise_mixture_rng == siss_randn + siss_add_noise_mixture, BIT-EXACT, all dtypes / layouts
  N shards with global offsets == 1 rank, BIT-EXACT
"""
import numpy as np
import pytest
import torch

from oracle import philox as P
from oracle import siss_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_device):
    from siss_b200 import _lib
    _lib.load()
    return cuda_device


@pytest.mark.parametrize("n,offset", [(4096, 0), (1000, 0), (1003, 437), (7, 2), (1 << 20, 1 << 33)])
def test_randn_matches_oracle_stream(n, offset, dev):
    from siss_b200.rng import DeviceRng
    rng = DeviceRng(seed=0x1234_5678_9ABC_DEF0)
    got = rng.randn((n,), torch.float32, dev, draw=5, elem_offset=offset).cpu().double().numpy()
    want = P.randn(n, seed=0x1234_5678_9ABC_DEF0, draw=5, elem_offset=offset)
    # MUFU lg2 / sin / cos: absolute error ~2^-21 on the unit-scale factors, radius <= 6.8. Where u1 -> 1 the radius
    # sqrt(-2 ln u1) is tiny and the log's ABSOLUTE error is amplified by 1/radius (a few 1e-5 on samples of size
    # ~1e-2, a handful per million) — hence a loose per-element bound and a tight bound on the mean error.
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)
    assert np.abs(got - want).mean() < 2e-6 and (np.abs(got - want) > 2e-5).mean() < 1e-4


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_randn_16bit_is_the_rounded_fp32_stream(dtype, dev):
    from siss_b200.rng import DeviceRng
    rng = DeviceRng(seed=99)
    for n, off in [(8192, 0), (1001, 4), (1001, 3)]:          # vector path, ragged tail, unaligned offset (scalar path)
        f32 = rng.randn((n,), torch.float32, dev, draw=1, elem_offset=off)
        assert torch.equal(rng.randn((n,), dtype, dev, draw=1, elem_offset=off), f32.to(dtype))


def test_randn_statistics_and_independence(dev):
    from siss_b200.rng import DeviceRng
    rng = DeviceRng(seed=7)
    z = rng.randn((64, 3, 256, 256), torch.float32, dev, draw=0)
    assert abs(z.mean().item()) < 1e-3 and abs(z.std().item() - 1) < 1e-3 and z.abs().max().item() < 6.8
    assert abs((z ** 4).mean().item() - 3) < 1e-2
    z2 = rng.randn((64, 3, 256, 256), torch.float32, dev, draw=1)
    assert abs((z * z2).mean().item()) < 1e-3 and not torch.equal(z, z2)
    assert torch.equal(z, rng.randn((64, 3, 256, 256), torch.float32, dev, draw=0))        # replayable


def test_draw_rows_matches_oracle_exactly(dev):
    from siss_b200.rng import DeviceRng
    for B, off, (lo, hi), lam in [(64, 0, (0, 1000), 0.5), (1000, 12345, (300, 1000), 0.25), (5, 2 ** 40, (999, 1000), 0.9)]:
        rng = DeviceRng(seed=46, row_offset=off)
        ts, keep = rng.draw_rows(B, dev, t_range=(lo, hi), lambd=lam, draw=9)
        want_t, want_k = P.draw_rows(B, seed=46, draw=9, t_lo=lo, t_hi=hi, lambd=lam, row_offset=off)
        assert ts.dtype == torch.int64 and keep.dtype == torch.uint8
        assert np.array_equal(ts.cpu().numpy(), want_t) and np.array_equal(keep.cpu().numpy().astype(bool), want_k)
    rng = DeviceRng(seed=46)
    assert rng.draw_rows(100, dev, lambd=1.0)[1].sum().item() == 0 and rng.draw_rows(100, dev, lambd=0.0)[1].sum().item() == 100
    ts, none = rng.draw_rows(10, dev, t_range=(0, 1000))
    assert none is None and ts.shape == (10,)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(4, 3, 256, 256), (64, 1, 28, 28), (7, 3, 33, 5), (3, 4, 64, 64), (2, 3, 512, 512)])
def test_fused_rng_mixture_is_bitwise_randn_then_mixture(dtype, shape, dev, monkeypatch):
    """TMA path here; the LDG path through the odd-D shape (scalar kernel) and through the SISS_NO_TMA=1 re-run of this
    file in a subprocess (tests/test_fullsize_gpu.py::test_ldg_kernels_pass_the_parity_suite)."""
    from siss_b200 import ops
    from siss_b200.rng import DeviceRng
    B = shape[0]
    torch.manual_seed(B)
    ac = O.make_alphas_cumprod(); g, s = O.gamma_sigma(ac)
    x0 = (torch.rand(shape) * 2 - 1).to(dtype).to(dev); a0 = (torch.rand(shape) * 2 - 1).to(dtype).to(dev)
    ts = torch.randint(0, 1000, (B,), device=dev); keep = (torch.rand(B) > 0.5).to(dev)
    per_row = x0[0].numel()
    for row_offset in (0, 5):
        rng = DeviceRng(seed=2024, row_offset=row_offset)
        noise = rng.randn(shape, dtype, dev, draw=3)
        want = ops.add_noise_mixture(x0, a0, noise, keep, ts, ac, g, s, 0.5)
        got = ops.add_noise_mixture_rng(x0, a0, keep, ts, ac, g, s, 0.5, 2024, 3, elem_offset=row_offset * per_row,
                                        want_noise=True)
        assert torch.equal(got[5], noise)
        for a, b, what in zip(got[:5], want, ("x_mix", "dist_x", "dist_a", "w_x", "w_a")):
            assert torch.equal(a, b), what
        assert got[0].dtype == dtype
        got2 = ops.add_noise_mixture_rng(x0, a0, keep, ts, ac, g, s, 0.5, 2024, 3, elem_offset=row_offset * per_row)
        assert got2[5] is None and torch.equal(got2[0], want[0]) and torch.equal(got2[3], want[3])


def test_sharded_draws_equal_one_rank(dev):
    """Two 'ranks' with global row offsets reproduce the 1-rank tensors — the property that removes the reference's
    same-draws-on-every-rank defect without any communication."""
    from siss_b200 import ops
    from siss_b200.rng import DeviceRng
    shape = (8, 3, 32, 32)
    ac = O.make_alphas_cumprod(); g, s = O.gamma_sigma(ac)
    torch.manual_seed(0)
    x0 = (torch.rand(shape) * 2 - 1).to(dev); a0 = (torch.rand(shape) * 2 - 1).to(dev)
    one = DeviceRng(seed=11)
    ts, keep = one.draw_rows(8, dev, t_range=(0, 1000), lambd=0.5, draw=0)
    full = ops.add_noise_mixture_rng(x0, a0, keep, ts, ac, g, s, 0.5, 11, 0)
    for r, (lo, hi) in enumerate([(0, 3), (3, 8)]):                   # ragged shards
        shard = DeviceRng(seed=11, row_offset=lo)
        ts_r, keep_r = shard.draw_rows(hi - lo, dev, t_range=(0, 1000), lambd=0.5, draw=0)
        assert torch.equal(ts_r, ts[lo:hi]) and torch.equal(keep_r, keep[lo:hi])
        part = ops.add_noise_mixture_rng(x0[lo:hi], a0[lo:hi], keep_r, ts_r, ac, g, s, 0.5, 11, 0,
                                         elem_offset=lo * x0[0].numel())
        assert torch.equal(part[0], full[0][lo:hi]) and torch.equal(part[3], full[3][lo:hi])


@pytest.mark.parametrize("loss_fn,kw", [("importance_sampling_with_mixture", dict(lambd=0.5, scaling_norm=5.0)),
                                        ("double_forward_with_neg_del", dict(scaling_norm=5.0)),
                                        ("naive_del", dict())])
def test_unlearn_step_with_device_rng_equals_explicit_draws(loss_fn, kw, dev):
    """UnlearnStep(device_rng=...) with nothing passed == the same step fed the stream's tensors explicitly."""
    import copy
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.rng import DeviceRng
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.manual_seed(1)
    net_a = torch.nn.Conv2d(1, 1, 3, padding=1).to(dev)
    net_b = copy.deepcopy(net_a)
    unet = lambda net: (lambda x, t, return_dict=False, **k: (net(x.float()),))
    sched = SissDDPMScheduler()
    B, shape = 6, (6, 1, 16, 16)
    x0 = (torch.rand(shape, device=dev) * 2 - 1); a0 = (torch.rand(shape, device=dev) * 2 - 1)
    ca, cb_ = GradCombiner(net_a.parameters()), GradCombiner(net_b.parameters())
    sa = UnlearnStep(unet(net_a), sched, ca, loss_fn=loss_fn, train_batch_size=B, max_norm=1.0,
                     device_rng=DeviceRng(seed=5), t_range=(300, 1000), **kw)
    sb = UnlearnStep(unet(net_b), sched, cb_, loss_fn=loss_fn, train_batch_size=B, max_norm=1.0, **kw)
    ref_rng = DeviceRng(seed=5)
    for it in range(2):
        out = sa.micro_step(x0, a0)
        ts, keep = ref_rng.draw_rows(B, dev, t_range=(300, 1000), lambd=0.5, draw=it)
        noise = ref_rng.randn(shape, torch.float32, dev, draw=it)
        assert torch.equal(out["timesteps"], ts) and int(ts.min()) >= 300
        sb.micro_step(x0, a0, noise, ts, keep_mask=keep if loss_fn.startswith("importance") else None)
        st_a, st_b = sa.sync_step(), sb.sync_step()
        assert torch.equal(st_a, st_b)
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            assert torch.equal(p.grad, q.grad)
    with pytest.raises(ValueError):
        sb.micro_step(x0, a0)                                      # no device_rng and no noise


@pytest.mark.parametrize("pred_dtype,tgt_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                   (torch.bfloat16, torch.bfloat16)])
@pytest.mark.parametrize("shape", [(4, 3, 64, 64), (5, 3, 17, 9), (2, 3, 256, 256)])
def test_erasediff_in_kernel_target(pred_dtype, tgt_dtype, shape, dev):
    """siss_dual_mse_rng_fwd_bwd: (a) the drawn target equals the oracle's aux stream EXACTLY (24-bit uniforms, rounded
    to the prediction dtype); (b) gradients / row sums are bit-identical to the existing dual-MSE kernel fed that target."""
    from siss_b200 import ops
    from siss_b200.rng import DeviceRng
    torch.manual_seed(shape[0])
    B = shape[0]; n = B * shape[1] * shape[2] * shape[3]
    px = torch.randn(shape).to(pred_dtype).to(dev); pa = torch.randn(shape).to(pred_dtype).to(dev)
    noise = torch.randn(shape).to(tgt_dtype).to(dev)
    for row_offset in (0, 3):
        off = row_offset * (n // B)
        gx, ga, rlx, rla, tgt = ops.dual_mse_rng_fwd_bwd(px, pa, noise, 0.25, 0.5, 99, 4, elem_offset=off, want_target=True)
        want_t = torch.from_numpy(P.rand_aux(n, seed=99, draw=4, elem_offset=off)).reshape(shape).to(pred_dtype)
        assert tgt.dtype == pred_dtype and torch.equal(tgt.cpu(), want_t)
        assert torch.equal(DeviceRng(99, row_offset=row_offset).rand_aux(shape, pred_dtype, dev, draw=4), tgt)
        tx = noise if tgt_dtype == pred_dtype else noise.to(pred_dtype)           # the unfused path needs one dtype
        rgx, rga, rrlx, rrla = ops.dual_mse_fwd_bwd(px, pa, tx, tgt, 0.25, 0.5)
        assert torch.equal(gx, rgx) and torch.equal(ga, rga)
        torch.testing.assert_close(rlx, rrlx, rtol=1e-5, atol=0); torch.testing.assert_close(rla, rrla, rtol=1e-5, atol=0)
        g2 = ops.dual_mse_rng_fwd_bwd(px, pa, noise, 0.25, 0.5, 99, 4, elem_offset=off)
        assert g2[4] is None and torch.equal(g2[1], ga)


def test_unlearn_step_erasediff_with_device_rng(dev):
    """UnlearnStep(loss_fn="erasediff", device_rng=...) == the same step fed the stream's eps / t / uniform target."""
    import copy
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.rng import DeviceRng
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.manual_seed(2)
    net_a = torch.nn.Conv2d(1, 1, 3, padding=1).to(dev); net_b = copy.deepcopy(net_a)
    unet = lambda net: (lambda x, t, return_dict=False, **k: (net(x.float()),))
    B, shape = 6, (6, 1, 16, 16)
    x0 = torch.rand(shape, device=dev) * 2 - 1; a0 = torch.rand(shape, device=dev) * 2 - 1
    mk = lambda net, rng: UnlearnStep(unet(net), SissDDPMScheduler(), GradCombiner(net.parameters()), loss_fn="erasediff",
                                      train_batch_size=B, eta=0.05, max_norm=1.0, device_rng=rng)
    sa, sb = mk(net_a, DeviceRng(seed=8)), mk(net_b, None)
    ref = DeviceRng(seed=8)
    for it in range(2):
        out = sa.micro_step(x0, a0)
        ts, _ = ref.draw_rows(B, dev, t_range=(0, 1000), draw=it)
        noise = ref.randn(shape, torch.float32, dev, draw=it)
        target = ref.rand_aux(shape, torch.float32, dev, draw=it)
        assert torch.equal(out["timesteps"], ts) and 0.0 <= float(target.min()) and float(target.max()) < 1.0
        sb.micro_step(x0, a0, noise, ts, forget_target=target)
        assert torch.equal(sa.sync_step(), sb.sync_step())
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            assert torch.equal(p.grad, q.grad) and p.grad.abs().sum().item() > 0
