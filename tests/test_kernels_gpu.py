"""GPU parity tests proper: every kernel is called through the C ABI (siss_b200.ops -> ctypes ->
libsiss_b200.so) and compared with (a) the golden fixtures produced by the reference's own loss
class and (b) the CPU oracle on the same seeded inputs.

Tolerances (also stated in DESIGN.md):
  * element-wise outputs (x_t, x_mix, loss_x/a, weighted losses, gradients into eps_hat): BIT-EXACT in
    fp32 / bf16 / fp16 — the kernels follow eager's op and rounding order;
  * per-sample sums d_x, d_a, row losses: rtol 1e-5 (fp32 reduction order);
  * importance weights: rtol = 8 eps32 (d_x + d_a) + 1e-5 — the reference subtracts two ~D/2-sized fp32
    sums before exp(), so its own weights carry that much summation noise; against the float64 oracle
    the kernel (which accumulates the difference directly) must be at least as close as the fp32
    reference is. Saturated weights (0, 1/(1-l), 1/l) must match exactly;
  * global norms and combined gradients: rtol 1e-5.
"""
import numpy as np
import pytest
import torch

from helpers import GOLDEN_CASES, conditioning_of, load_golden, weight_tolerance
from oracle import siss_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_device):
    from siss_b200 import _lib
    _lib.load()
    return cuda_device


def _bits_equal(a: torch.Tensor, b: torch.Tensor, what=""):
    a, b = a.detach().cpu(), b.detach().cpu()
    assert a.dtype == b.dtype and a.shape == b.shape, f"{what}: {a.dtype}{tuple(a.shape)} vs {b.dtype}{tuple(b.shape)}"
    if not torch.equal(a.float().nan_to_num(nan=-7.0), b.float().nan_to_num(nan=-7.0)):
        d = (a.double() - b.double()).abs()
        raise AssertionError(f"{what}: not bit-exact, max abs diff {d.max().item():.3e} at {d.argmax().item()}, "
                             f"{(d > 0).sum().item()} / {d.numel()} differ")


def _gpu_case(name, dev):
    c = load_golden(name)
    g = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in c.items()}
    return c, g


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_add_noise_bit_exact(name, dev):
    from siss_b200 import ops
    c, g = _gpu_case(name, dev)
    _bits_equal(ops.add_noise(g["x0"], g["noise"], g["t"], g["alphas_cumprod"]), c["xt_x"], "add_noise keep")
    xt_x, xt_a = ops.add_noise_pair(g["x0"], g["a0"], g["noise"], g["t"], g["alphas_cumprod"])
    _bits_equal(xt_x, c["xt_x"], "pair keep")
    _bits_equal(xt_a, c["xt_a"], "pair forget")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(3, 1, 28, 28), (5, 3, 17, 9), (2, 4, 64, 64), (1, 3, 256, 256), (130, 1, 8, 8)])
def test_add_noise_vs_oracle_shapes(dtype, shape, dev):
    """ragged / odd D (scalar path), rows longer than one tile, more rows than CTAs, every timestep edge."""
    from siss_b200.scheduler import SissDDPMScheduler
    torch.manual_seed(sum(shape) * 131 + {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}[dtype])
    sched = SissDDPMScheduler()
    x0 = (torch.rand(shape) * 2 - 1).to(dtype)
    a0 = (torch.rand(shape) * 2 - 1).to(dtype)
    n = torch.randn(shape).to(dtype)
    t = torch.randint(0, 1000, (shape[0],))
    t[0], t[-1] = 0, 999
    ex = O.add_noise(sched.alphas_cumprod, x0, n, t)
    ea = O.add_noise(sched.alphas_cumprod, a0, n, t)
    _bits_equal(sched.add_noise(x0.to(dev), n.to(dev), t.to(dev)), ex, "single")
    gx, ga = sched.add_noise_pair(x0.to(dev), a0.to(dev), n.to(dev), t.to(dev))
    _bits_equal(gx, ex, "pair x"); _bits_equal(ga, ea, "pair a")


def test_add_noise_empty_and_misaligned(dev):
    from siss_b200 import ops
    sched_ac = O.make_alphas_cumprod().to(dev)
    e = torch.empty(0, 1, 4, 4, device=dev)
    assert ops.add_noise(e, e, torch.empty(0, dtype=torch.long, device=dev), sched_ac).shape == e.shape
    # a view that starts 4 bytes into an allocation: 16B alignment fails -> scalar path, same answer
    torch.manual_seed(3)
    base = torch.rand(2 * 16 + 1, device=dev)
    x = base[1:].view(2, 1, 4, 4)
    n = torch.randn(2 * 16 + 1, device=dev)[1:].view(2, 1, 4, 4)
    t = torch.tensor([10, 900], device=dev)
    assert x.data_ptr() % 16 != 0 and x.is_contiguous()
    _bits_equal(ops.add_noise(x, n, t, sched_ac), O.add_noise(sched_ac.cpu(), x.cpu(), n.cpu(), t.cpu()), "misaligned")


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def _check_weights(w_x, w_a, c, dist_x, dist_a):
    """kernel weights vs the reference's golden fp32 weights with the conditioning-derived tolerance."""
    gw_x, gw_a = c["siss_w_x"].double(), c["siss_w_a"].double()
    tol = weight_tolerance(dist_x.cpu(), dist_a.cpu())
    for got, exp, nm in ((w_x.cpu().double(), gw_x, "w_x"), (w_a.cpu().double(), gw_a, "w_a")):
        sat = (exp == 0) | torch.isinf(exp)
        assert torch.equal(got[sat], exp[sat]), f"{nm}: saturated entries must match exactly: {got} vs {exp}"
        rel = ((got - exp).abs() / exp.abs().clamp_min(1e-300))[~sat]
        assert (rel <= tol[~sat]).all(), f"{nm}: rel err {rel} > tol {tol[~sat]}"


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mixture_weights_golden(name, dev):
    from siss_b200 import ops
    c, g = _gpu_case(name, dev)
    x_mix, d_x, d_a, w_x, w_a = ops.mixture_weights(g["xt_x"], g["xt_a"], g["x0"], g["a0"], g["siss_keep_mask"],
                                                    g["t"], g["gamma"], g["sigma"], c["lambd"])
    exp_mix = O.select_mixture(c["xt_x"], c["xt_a"], c["siss_keep_mask"])
    _bits_equal(x_mix, exp_mix, "x_mix")
    ed_x, ed_a = O.gaussian_exponents(exp_mix, c["x0"], c["a0"], c["gamma"][c["t"]], c["sigma"][c["t"]])
    torch.testing.assert_close(d_x.cpu(), ed_x, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(d_a.cpu(), ed_a, rtol=1e-5, atol=1e-6)
    _check_weights(w_x, w_a, c, d_x, d_a)
    # fused K1 o K2 gives the same x_mix bit for bit and the same row scalars
    f_mix, fd_x, fd_a, fw_x, fw_a = ops.add_noise_mixture(g["x0"], g["a0"], g["noise"], g["siss_keep_mask"], g["t"],
                                                           g["alphas_cumprod"], g["gamma"], g["sigma"], c["lambd"])
    _bits_equal(f_mix, exp_mix, "fused x_mix")
    for a, b in ((fd_x, d_x), (fd_a, d_a), (fw_x, w_x), (fw_a, w_a)):
        _bits_equal(a, b, "fused row scalars")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_weights_vs_float64_oracle(name, dev):
    """Against the float64 evaluation of the same formula the kernel must be at least as accurate as
    the reference's own fp32 result (it accumulates d_x - d_a directly)."""
    from siss_b200 import ops
    c, g = _gpu_case(name, dev)
    _, d_x, d_a, w_x, w_a = ops.mixture_weights(g["xt_x"], g["xt_a"], g["x0"], g["a0"], g["siss_keep_mask"],
                                                g["t"], g["gamma"], g["sigma"], c["lambd"])
    mix = O.select_mixture(c["xt_x"], c["xt_a"], c["siss_keep_mask"]).double()
    g64, s64 = c["gamma"].double()[c["t"]], c["sigma"].double()[c["t"]]
    d64x, d64a = O.gaussian_exponents(mix, c["x0"].double(), c["a0"].double(), g64, s64)
    e_wx, e_wa = O.importance_weights(d64x, d64a, c["lambd"])
    for got, ref32, exact in ((w_x, c["siss_w_x"], e_wx), (w_a, c["siss_w_a"], e_wa)):
        got, ref32 = got.cpu().double(), ref32.double()
        ok = torch.isfinite(exact) & (exact > 1e-30) & (exact < 1e30)
        err_k = ((got - exact).abs() / exact)[ok]
        err_r = ((ref32 - exact).abs() / exact)[ok]
        assert (err_k <= err_r + 2e-5).all(), f"kernel err {err_k} vs reference fp32 err {err_r}"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,lambd", [((4, 3, 256, 256), 0.5), ((64, 1, 28, 28), 0.5), ((7, 3, 33, 5), 0.25),
                                         ((300, 4, 8, 8), 0.5), ((2, 4, 64, 64), 0.9)])
def test_mixture_vs_oracle_shapes(dtype, shape, lambd, dev):
    """multi-tile rows (cross-CTA ticket reduce), B > grid, odd D, t in {0, 999} and t == 999 for all."""
    from siss_b200 import ops
    torch.manual_seed(shape[0] * 7 + int(lambd * 100))
    ac = O.make_alphas_cumprod(); gamma, sigma = O.gamma_sigma(ac)
    x0 = (torch.rand(shape) * 2 - 1).to(dtype); a0 = (torch.rand(shape) * 2 - 1).to(dtype)
    n = torch.randn(shape).to(dtype)
    t = torch.full((shape[0],), 999)
    if shape[0] > 4:
        t = torch.randint(0, 1000, (shape[0],)); t[0] = 0; t[1] = 999
    keep = torch.rand(shape[0]) > lambd
    xt_x, xt_a = O.add_noise(ac, x0, n, t), O.add_noise(ac, a0, n, t)
    mix = O.select_mixture(xt_x, xt_a, keep)
    ed_x, ed_a = O.gaussian_exponents(mix, x0, a0, gamma[t], sigma[t])
    ew_x, ew_a = O.importance_weights(ed_x, ed_a, lambd)
    out = ops.add_noise_mixture(x0.to(dev), a0.to(dev), n.to(dev), keep, t.to(dev), ac, gamma, sigma, lambd)
    _bits_equal(out[0], mix, "x_mix")
    torch.testing.assert_close(out[1].cpu(), ed_x, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(out[2].cpu(), ed_a, rtol=1e-5, atol=1e-6)
    fake = {"siss_w_x": ew_x, "siss_w_a": ew_a}
    _check_weights(out[3], out[4], fake, out[1], out[2])
    # identity of the two weights: (1-l) w_x + l w_a == 1 whenever neither saturates to inf/nan
    ident = (1 - lambd) * out[3].double() + lambd * out[4].double()
    torch.testing.assert_close(ident.cpu(), torch.ones_like(ident.cpu()), rtol=1e-6, atol=1e-6)
    # workspace is left clean: run again, identical bits
    out2 = ops.add_noise_mixture(x0.to(dev), a0.to(dev), n.to(dev), keep, t.to(dev), ac, gamma, sigma, lambd)
    for a, b in zip(out, out2):
        _bits_equal(a, b, "rerun determinism")


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_wmse_golden(name, dev):
    """With the reference's own weights as input, every element-wise output of K3 is bit-exact."""
    from siss_b200 import ops
    c, g = _gpu_case(name, dev)
    mix = O.select_mixture(c["xt_x"], c["xt_a"], c["siss_keep_mask"]).to(dev)
    args = (g["siss_pred"], mix, g["x0"], g["a0"], g["t"], g["gamma"], g["sigma"], g["siss_w_x"], g["siss_w_a"])
    lx, la, wlx, wla = ops.wmse_fwd(*args)
    _bits_equal(lx, c["siss_loss_x"], "loss_x"); _bits_equal(la, c["siss_loss_a"], "loss_a")
    _bits_equal(wlx, c["siss_wl_x"], "weighted_loss_x"); _bits_equal(wla, c["siss_wl_a"], "weighted_loss_a")
    B = c["x0"].shape[0]
    go = float(np.float32(1.0) / np.float32(B))
    gx, ga, rlx, rla = ops.wmse_fwd_bwd(*args, go, go)
    _bits_equal(gx, c["siss_grad_x"], "grad_x"); _bits_equal(ga, c["siss_grad_a"], "grad_a")
    torch.testing.assert_close(rlx.cpu(), c["siss_loss_x"].sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(rla.cpu(), c["siss_loss_a"].sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-6)
    # general backward: scalar-broadcast upstream (what .sum() gives) and dense upstream
    go_t = torch.full((), go, device=dev).expand(lx.shape)
    _bits_equal(ops.wmse_bwd(*args, go_wloss_x=go_t), c["siss_grad_x"], "bwd scalar x")
    _bits_equal(ops.wmse_bwd(*args, go_wloss_a=go_t.contiguous()), c["siss_grad_a"], "bwd dense a")


@pytest.mark.parametrize("pred_dtype,in_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                                 (torch.float32, torch.float16), (torch.bfloat16, torch.bfloat16),
                                                 (torch.float16, torch.float16)])
@pytest.mark.parametrize("shape", [(2, 3, 256, 256), (9, 1, 28, 28), (3, 3, 7, 5)])
def test_wmse_vs_autograd_oracle(pred_dtype, in_dtype, shape, dev):
    """K3 against eager autograd on CPU for all supported (pred, latent) dtype pairs."""
    from siss_b200 import ops
    torch.manual_seed(shape[0] * 31 + shape[-1])
    ac = O.make_alphas_cumprod(); gamma, sigma = O.gamma_sigma(ac)
    B = shape[0]
    x0 = (torch.rand(shape) * 2 - 1).to(in_dtype); a0 = (torch.rand(shape) * 2 - 1).to(in_dtype)
    mix = torch.randn(shape).to(in_dtype)
    t = torch.randint(100, 1000, (B,)); t[0] = 999
    w_x, w_a = torch.rand(B) * 2, torch.rand(B) * 2
    pred = torch.randn(shape).to(pred_dtype).requires_grad_(True)
    g_t, s_t = gamma[t].view(-1, 1, 1, 1), sigma[t].view(-1, 1, 1, 1)
    lx = (pred - (mix - g_t * x0) / s_t) ** 2
    la = (pred - (mix - g_t * a0) / s_t) ** 2
    wlx, wla = w_x.view(-1, 1, 1, 1) * lx, w_a.view(-1, 1, 1, 1) * la
    G, Bcfg = 4, 8
    ((wlx.sum() / Bcfg) / G).backward(retain_graph=True); egx = pred.grad.clone(); pred.grad = None
    ((wla.sum() / Bcfg) / G).backward(); ega = pred.grad.clone()
    from siss_b200.step import upstream_scale
    go = upstream_scale(Bcfg, G)
    d = lambda v: v.detach().to(dev)
    args = (d(pred), d(mix), d(x0), d(a0), d(t), gamma, sigma, d(w_x), d(w_a))
    o = ops.wmse_fwd(*args)
    for got, exp, nm in zip(o, (lx, la, wlx, wla), ("lx", "la", "wlx", "wla")):
        _bits_equal(got, exp.float(), nm)
    gx, ga, rlx, rla = ops.wmse_fwd_bwd(*args, go, go)
    _bits_equal(gx, egx, "grad_x"); _bits_equal(ga, ega, "grad_a")
    torch.testing.assert_close(rlx.cpu(), lx.float().sum(dim=[1, 2, 3]).detach(), rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# squared-error family
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pd,td", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                   (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16),
                                   (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("n", [(2, 1, 28, 28), (3, 3, 5, 7)])
def test_sqerr_fwd_bwd_vs_autograd(pd, td, n, dev):
    from siss_b200 import ops
    torch.manual_seed(17)
    pred = torch.randn(n).to(pd).requires_grad_(True); tgt = torch.randn(n).to(td)
    loss = (pred - tgt) ** 2
    scaled = -1.7 * loss
    l, s = ops.sqerr_fwd(pred.detach().to(dev), tgt.to(dev), alpha=-1.7)
    _bits_equal(l, loss.detach(), "loss"); _bits_equal(s, scaled.detach(), "scaled")
    (scaled.sum() / 4).backward(retain_graph=True); e1 = pred.grad.clone(); pred.grad = None
    (loss.sum() / 4).backward(); e2 = pred.grad.clone()
    go = torch.full((), 0.25, dtype=loss.dtype, device=dev).expand(n)
    _bits_equal(ops.sqerr_bwd(pred.detach().to(dev), tgt.to(dev), go_scaled=go, alpha=-1.7), e1, "bwd scaled")
    _bits_equal(ops.sqerr_bwd(pred.detach().to(dev), tgt.to(dev), go_loss=go), e2, "bwd loss")


@pytest.mark.parametrize("shape", [(4, 3, 64, 64), (3, 1, 7, 3), (2, 3, 256, 256)])
@pytest.mark.parametrize("td", [torch.float32, torch.bfloat16])
def test_dual_mse(shape, td, dev):
    from siss_b200 import ops
    torch.manual_seed(23)
    px, pa = torch.randn(shape), torch.randn(shape)
    tx, ta = torch.randn(shape).to(td), torch.rand(shape).to(td)
    go = 1.0 / 64
    for tgt_a in (tx, ta):  # shared target (No-IS) and separate target (EraseDiff)
        gx, ga, rx, ra = ops.dual_mse_fwd_bwd(px.to(dev), pa.to(dev), tx.to(dev), tx.to(dev) if tgt_a is tx else ta.to(dev),
                                              go, go)
        go32 = torch.tensor(go, dtype=torch.float32)
        _bits_equal(gx, go32 * (2 * (px - tx)), "gx")
        _bits_equal(ga, go32 * (2 * (pa - tgt_a)), "ga")
        torch.testing.assert_close(rx.cpu(), ((px - tx) ** 2).sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(ra.cpu(), ((pa - tgt_a) ** 2).sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 3, 4, 1023, 4096 * 5 + 1, (1 << 21) + 7])
@pytest.mark.parametrize("mode", ["scaling_norm", "erasediff"])
def test_norm3_combine_vs_oracle(n, mode, dev):
    """K4a/K4b vs the oracle evaluated in float64 (tight) and in float32 (loose: the CPU reference's
    own fp32 norm / dot reductions are only good to ~1e-5..1e-4 at 2M elements, measured)."""
    from siss_b200 import ops, _lib
    torch.manual_seed(n % 1000 + 1)
    gx, ga = torch.randn(n) * 3e-2, torch.randn(n) * 1e-2 + 2e-3
    kw = dict(scaling_norm=5.0) if mode == "scaling_norm" else dict(eta=0.05)
    sums = ops.norm3(gx.to(dev), ga.to(dev))
    ref = torch.tensor([(gx.double() ** 2).sum(), (ga.double() ** 2).sum(), (gx.double() * ga.double()).sum()])
    torch.testing.assert_close(sums.cpu(), ref, rtol=1e-12, atol=1e-300)     # exact fp64 products
    m = _lib.SISS_COMBINE_SCALING_NORM if mode == "scaling_norm" else _lib.SISS_COMBINE_ERASEDIFF
    out, stats = ops.combine(gx.to(dev), ga.to(dev), sums, m, 5.0 if mode == "scaling_norm" else 0.05, 1.0)
    st = stats.cpu().double()
    for dtype, rtol in ((torch.float64, 2e-6), (torch.float32, 2e-4)):
        exp, nx, na, s, tn, clip = O.combine_flat(gx.to(dtype), ga.to(dtype), max_norm=1.0, **kw)
        scale = float(exp.abs().max())
        torch.testing.assert_close(st[0], nx.double(), rtol=rtol, atol=0)
        torch.testing.assert_close(st[1], na.double(), rtol=rtol, atol=0)
        torch.testing.assert_close(st[2], s.double(), rtol=rtol, atol=rtol * 0.1)
        torch.testing.assert_close(st[3], tn.double(), rtol=rtol, atol=1e-8)
        torch.testing.assert_close(st[4], clip.double(), rtol=rtol, atol=0)
        # x - s a cancels element-wise where x ~ s a (and s itself is an fp32 value): absolute tolerance
        # relative to the tensor's scale and to the magnitude of the two terms being subtracted
        mag = float((gx.double().abs() + (s.double() * ga.double()).abs()).max())
        torch.testing.assert_close(out.cpu().double(), exp.double(), rtol=rtol, atol=max(rtol * scale, 4e-7 * mag))


def test_combine_inplace_misaligned_and_guards(dev):
    from siss_b200 import ops, _lib
    torch.manual_seed(2)
    n = 10007
    base_x, base_a = torch.randn(n + 1, device=dev), torch.randn(n + 1, device=dev)
    gx, ga = base_x[1:], base_a[1:]            # 4-byte offset: scalar path
    exp, *_ = O.combine_flat(gx.cpu(), ga.cpu(), scaling_norm=500.0, max_norm=1.0)
    sums = ops.norm3(gx, ga)
    out, _ = ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0, out=gx)  # aliasing g_x
    assert out.data_ptr() == gx.data_ptr()
    torch.testing.assert_close(gx.cpu(), exp, rtol=1e-4, atol=1e-4 * float(exp.abs().max()))
    # zero NegGrad gradient: s = inf without the guard (celeb/sd), 0 with it (tshirt)
    z = torch.zeros(64, device=dev); x = torch.ones(64, device=dev) * 0.01
    sums = ops.norm3(x, z)
    _, st = ops.combine(x, z, sums, _lib.SISS_COMBINE_SCALING_NORM, 5.0, 1.0, inf_guard=True)
    assert st[2].item() == 0.0
    o, st = ops.combine(x, z, sums, _lib.SISS_COMBINE_SCALING_NORM, 5.0, 1.0, inf_guard=False)
    assert torch.isinf(st[2]).item() and torch.isnan(o).all().item()       # inf * 0 = nan, as in eager
    # no clip requested
    o, st = ops.combine(x * 1e3, z, ops.norm3(x * 1e3, z), _lib.SISS_COMBINE_NONE, 0.0, 0.0)
    assert st[4].item() == 1.0 and torch.equal(o, x * 1e3)


def test_workspace_reuse_across_batch_sizes(dev):
    """One zero-initialised row workspace serves calls with different B and D in any order (its layout
    does not depend on the per-call B). Regression test: partial sums of a small-B call used to land
    where a later large-B call kept its ticket / span-claim counters."""
    from siss_b200 import ops
    ac = O.make_alphas_cumprod(); gamma, sigma = O.gamma_sigma(ac)
    order = [(4, 3, 256, 256), (64, 1, 28, 28), (300, 4, 8, 8), (2, 3, 256, 256), (130, 1, 28, 28), (4, 3, 256, 256),
             (64, 1, 28, 28)]
    for i, shape in enumerate(order):
        torch.manual_seed(100 + i)
        B = shape[0]
        x0 = torch.rand(shape) * 2 - 1; a0 = torch.rand(shape) * 2 - 1; n = torch.randn(shape)
        t = torch.randint(0, 1000, (B,)); keep = torch.rand(B) > 0.5
        mix = O.select_mixture(O.add_noise(ac, x0, n, t), O.add_noise(ac, a0, n, t), keep)
        ed_x, ed_a = O.gaussian_exponents(mix, x0, a0, gamma[t], sigma[t])
        out = ops.add_noise_mixture(x0.to(dev), a0.to(dev), n.to(dev), keep, t.to(dev), ac, gamma, sigma, 0.5)
        _bits_equal(out[0], mix, f"x_mix {shape}")
        torch.testing.assert_close(out[1].cpu(), ed_x, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out[2].cpu(), ed_a, rtol=1e-5, atol=1e-6)
        pred = torch.randn(shape)
        gx, ga, rlx, rla = ops.wmse_fwd_bwd(pred.to(dev), out[0], x0.to(dev), a0.to(dev), t.to(dev), gamma, sigma,
                                            out[3], out[4], 1 / 64, 1 / 64)
        eps_x = (mix - gamma[t].view(-1, 1, 1, 1) * x0) / sigma[t].view(-1, 1, 1, 1)
        torch.testing.assert_close(rlx.cpu(), ((pred - eps_x) ** 2).sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("mode", ["scaling_norm", "erasediff"])
def test_multi_tensor_combine_matches_reference_dict_loop(mode, dev):
    """siss_mt_norm3 / siss_mt_combine over a ragged list of per-parameter tensors (sizes 1 .. 70k, one of
    them a 4-byte-offset view so its chunks take the scalar path) vs the oracle's flat evaluation of the
    reference's dict loop (delete_celeb.py:717-767), and vs the flat-buffer kernels."""
    from siss_b200 import ops, _lib
    torch.manual_seed(4)
    sizes = [1, 3, 4096, 4097, 70001, 513, 8192, 7]
    xs = [torch.randn(n) * 3e-2 for n in sizes]
    as_ = [torch.randn(n) * 1e-2 + 2e-3 for n in sizes]
    dx = [t.to(dev) for t in xs]; da = [t.to(dev) for t in as_]
    base = torch.zeros(sizes[5] + 1, device=dev); base[1:] = dx[5]; dx[5] = base[1:]      # misaligned tensor
    assert dx[5].data_ptr() % 16 != 0
    outs = [torch.empty_like(t) for t in dx]
    plan = ops.MultiTensorPlan(dx, da, outs)
    kw = dict(scaling_norm=5.0) if mode == "scaling_norm" else dict(eta=0.05)
    m = _lib.SISS_COMBINE_SCALING_NORM if mode == "scaling_norm" else _lib.SISS_COMBINE_ERASEDIFF
    stats = plan.combine(m, 5.0 if mode == "scaling_norm" else 0.05, 1.0).cpu().double()
    fx, fa = torch.cat(xs), torch.cat(as_)
    ref = torch.tensor([(fx.double() ** 2).sum(), (fa.double() ** 2).sum(), (fx.double() * fa.double()).sum()])
    torch.testing.assert_close(plan.sums3.cpu(), ref, rtol=1e-12, atol=1e-300)
    exp, nx, na, s, tn, clip = O.combine_flat(fx.double(), fa.double(), max_norm=1.0, **kw)
    got = torch.cat([o.cpu() for o in outs]).double()
    scale = float(exp.abs().max())
    torch.testing.assert_close(got, exp, rtol=2e-6, atol=max(2e-6 * scale, 4e-7 * float((fx.abs() + (s * fa).abs()).max())))
    torch.testing.assert_close(stats, torch.stack([nx, na, s.double(), tn, clip.double()]), rtol=2e-6, atol=2e-7)
    # identical to the flat-buffer kernels on the concatenated tensors (same arithmetic, other reduction order)
    flat_out, flat_stats = ops.combine(fx.to(dev), fa.to(dev), ops.norm3(fx.to(dev), fa.to(dev)), m,
                                       5.0 if mode == "scaling_norm" else 0.05, 1.0)
    torch.testing.assert_close(got.float(), flat_out.cpu(), rtol=1e-6, atol=1e-9)
    # in-place form: out defaults to accum_x
    plan2 = ops.MultiTensorPlan(dx, da)
    plan2.combine(m, 5.0 if mode == "scaling_norm" else 0.05, 1.0)
    torch.testing.assert_close(torch.cat([t.cpu() for t in dx]).double(), got, rtol=0, atol=0)
