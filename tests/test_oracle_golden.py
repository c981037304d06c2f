"""CPU: pin the oracle against outputs of the reference's own loss class (tests/golden/*.npz were
produced by executing /root/reference/losses/ddpm_deletion_loss.py, see make_golden.py). The oracle
is a torch-CPU restatement, so for the loss class the bar is BIT-EXACT."""
import numpy as np
import pytest
import torch

from helpers import (GOLDEN_CASES, MEMBERSHIP_CASES, MembershipStubUNet, conditioning_of, load_golden,
                     load_membership_golden)
from oracle import siss_oracle as O


def _dicts(c):
    return ({"og_latents": c["x0"], "noisy_latents": c["xt_x"]},
            {"og_latents": c["a0"], "noisy_latents": c["xt_a"]})


def _eq(a: torch.Tensor, b: torch.Tensor, what: str):
    assert a.shape == b.shape, what
    assert torch.equal(a.float().nan_to_num(nan=-7.0), b.float().nan_to_num(nan=-7.0)), \
        f"{what}: max abs diff {(a.float() - b.float()).abs().max().item()}"


def test_golden_present():
    assert len(GOLDEN_CASES) >= 9


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_schedule_tables(name):
    c = load_golden(name)
    kw = dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012) if c["schedule"] == "scaled_linear" \
        else dict(beta_schedule="linear")
    ac = O.make_alphas_cumprod(**kw)
    _eq(ac, c["alphas_cumprod"], "alphas_cumprod")
    g, s = O.gamma_sigma(ac)
    _eq(g, c["gamma"], "gamma")
    _eq(s, c["sigma"], "sigma")


def test_schedule_known_values():
    # values computed during the survey
    ac = O.make_alphas_cumprod()
    g, s = O.gamma_sigma(ac)
    assert abs(g[999].item() - 0.00635) < 1e-5 and abs(s[999].item() - 0.99998) < 1e-5 and abs(s[0].item() - 0.0100) < 1e-4
    g_sd, _ = O.gamma_sigma(O.make_alphas_cumprod(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012))
    assert abs(g_sd[999].item() - 0.06826) < 1e-4


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_add_noise(name):
    c = load_golden(name)
    _eq(O.add_noise(c["alphas_cumprod"], c["x0"], c["noise"], c["t"]), c["xt_x"], "xt_x")
    _eq(O.add_noise(c["alphas_cumprod"], c["a0"], c["noise"], c["t"]), c["xt_a"], "xt_a")


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_siss_forward_and_grads(name):
    c = load_golden(name)
    all_d, del_d = _dicts(c)
    loss = O.OracleDeletionLoss(c["gamma"], c["sigma"])
    unet = O.StubUNet()
    preds = []
    orig = unet.forward

    def fwd(*a, **k):
        out = orig(*a, **k)
        out[0].retain_grad()
        preds.append(out[0])
        return out
    unet.forward = fwd
    torch.manual_seed(c["siss_mask_seed"])           # same seed, same CPU draw position as the reference
    items = loss.importance_sampling_with_mixture(unet, c["t"], c["noise"], conditioning_of(c), all_d, del_d,
                                                  lambd=c["lambd"])
    assert items[0] is None
    for k, v in zip(("loss_x", "loss_a", "w_x", "w_a", "wl_x", "wl_a"), items[1:]):
        _eq(v, c[f"siss_{k}"], k)
    _eq(preds[0], c["siss_pred"], "pred")
    B = c["x0"].shape[0]
    (items[5].sum() / B).backward(retain_graph=True)
    _eq(preds[0].grad, c["siss_grad_x"], "grad_x")
    preds[0].grad = None
    (items[6].sum() / B).backward()
    _eq(preds[0].grad, c["siss_grad_a"], "grad_a")
    # the explicit-mask entry point reproduces the draw
    torch.manual_seed(c["siss_mask_seed"])
    assert torch.equal(O.draw_keep_mask(B, c["lambd"]), c["siss_keep_mask"])


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_other_methods(name):
    c = load_golden(name)
    all_d, del_d = _dicts(c)
    cond = conditioning_of(c)
    loss = O.OracleDeletionLoss(c["gamma"], c["sigma"])
    it = loss.double_forward_with_neg_del(O.StubUNet(), c["t"], c["noise"], cond, all_d, del_d)
    _eq(it[1], c["nois_loss_x"], "nois_loss_x"); _eq(it[2], c["nois_loss_a"], "nois_loss_a")
    assert it[5] is it[1] and it[6] is it[2] and it[0] is None and it[3] is None

    torch.manual_seed(c["erasediff_seed"])
    it = loss.erasediff(O.StubUNet(), c["t"], c["noise"], cond, all_d, del_d)
    _eq(it[1], c["erasediff_loss_x"], "erasediff_loss_x"); _eq(it[2], c["erasediff_loss_a"], "erasediff_loss_a")

    it = loss.simple_neg_del(O.StubUNet(), c["t"], c["noise"], cond, all_d, del_d, superfactor=1.7)
    _eq(it[0], c["neg_loss"], "neg_loss"); _eq(it[2], c["neg_loss_a"], "neg_loss_a")
    assert it[1] is None

    it = loss.naive_del(O.StubUNet(), c["t"], c["noise"], cond, all_d, del_d)
    _eq(it[0], c["naive_loss"], "naive_loss")
    assert it[1] is it[0] and it[2] is None

    torch.manual_seed(c["subscore_seed"])
    if "subscore_raises" in c:
        with pytest.raises(ZeroDivisionError):
            loss.subscore_bernoulli(O.StubUNet(), c["t"], c["noise"], cond, all_d, del_d, lambd=c["lambd"])
    else:
        it = loss.subscore_bernoulli(O.StubUNet(), c["t"], c["noise"], cond, all_d, del_d, lambd=c["lambd"])
        _eq(it[1], c["subscore_loss_x"], "subscore_loss_x"); _eq(it[2], c["subscore_loss_a"], "subscore_loss_a")


def test_saturation_values_in_golden():
    """The reference's weights saturate to exactly 0, 1/(1-lambd), 1/lambd when exp overflows
    (SURVEY.md §7): make sure the fixtures really exercise that."""
    c = load_golden("small_t_saturate")
    lam = c["lambd"]
    wx, wa = c["siss_w_x"], c["siss_w_a"]
    assert (wx == 0).any() and (wa == 0).any()
    assert torch.isin(wx, torch.tensor([0.0, 1 / (1 - lam)], dtype=torch.float32)).all()
    assert torch.isin(wa, torch.tensor([0.0, 1 / lam], dtype=torch.float32)).all()


# ---------------------------------------------------------------------------------------------
# combine block: the reference has no function for it (inline in run()); cross-check the oracle's
# dict-based restatement against a literal flat evaluation and against plain autograd identities.
# ---------------------------------------------------------------------------------------------
class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.c1 = torch.nn.Conv2d(1, 4, 3, padding=1)
        self.c2 = torch.nn.Conv2d(4, 1, 3, padding=1)

    def forward(self, x, timesteps, return_dict=False, **kw):
        return (self.c2(torch.tanh(self.c1(x))),)


def _run_loop(loss_fn, G, scaling_norm=None, eta=None, inf_guard=False):
    torch.manual_seed(11)
    net = TinyNet()
    ac = O.make_alphas_cumprod()
    gamma, sigma = O.gamma_sigma(ac)
    loss = O.OracleDeletionLoss(gamma, sigma)
    B = 4
    loop = O.ReferenceGradLoop(net, train_batch_size=B, grad_accum_steps=G)
    flat_x = flat_a = None
    for k in range(G):
        x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
        noise = torch.randn(B, 1, 8, 8)
        t = torch.randint(300, 1000, (B,))
        all_d = {"og_latents": x0, "noisy_latents": O.add_noise(ac, x0, noise, t)}
        del_d = {"og_latents": a0, "noisy_latents": O.add_noise(ac, a0, noise, t)}
        kw = {"lambd": 0.5} if loss_fn == "importance_sampling_with_mixture" else {}
        items = getattr(loss, loss_fn)(net, t, noise, {}, all_d, del_d, **kw)
        # independent evaluation of the two accumulated gradients with autograd.grad
        params = list(net.parameters())
        gx = torch.autograd.grad(items[5].sum() / B / G, params, retain_graph=True)
        ga = torch.autograd.grad(items[6].sum() / B / G, params, retain_graph=True)
        fx = torch.cat([g.reshape(-1) for g in gx]); fa = torch.cat([g.reshape(-1) for g in ga])
        flat_x = fx if flat_x is None else flat_x + fx
        flat_a = fa if flat_a is None else flat_a + fa
        loop.micro_step(items, retain_graph=(loss_fn == "importance_sampling_with_mixture"))
    out = loop.sync_step(False, loss_fn, scaling_norm=scaling_norm, eta=eta, inf_guard=inf_guard)
    got = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    exp, nx, na, s, tn, clip = O.combine_flat(flat_x, flat_a, scaling_norm=scaling_norm, eta=eta, inf_guard=inf_guard)
    return got, exp, out, (nx, na, s, tn, clip)


@pytest.mark.parametrize("loss_fn,kw", [
    ("importance_sampling_with_mixture", dict(scaling_norm=5.0)),
    ("double_forward_with_neg_del", dict(scaling_norm=500.0)),
    ("erasediff", dict(eta=0.01)),
])
@pytest.mark.parametrize("G", [1, 3])
def test_combine_matches_literal_loop(loss_fn, kw, G):
    got, exp, out, (nx, na, s, tn, clip) = _run_loop(loss_fn, G, **kw)
    # the loop accumulates through clone/subtract, the flat form directly: same maths, different
    # fp32 rounding order -> rtol 1e-5 (the tolerance DESIGN.md states for combined gradients)
    torch.testing.assert_close(got, exp, rtol=2e-5, atol=1e-7)
    torch.testing.assert_close(out["norm_x"], nx, rtol=1e-5, atol=0)
    torch.testing.assert_close(out["norm_a"], na, rtol=1e-5, atol=0)
    torch.testing.assert_close(out["scaling_factor"].float(), s.float(), rtol=1e-4, atol=1e-7)
    assert torch.linalg.vector_norm(got) <= 1.0 + 1e-5


@pytest.mark.parametrize("name", MEMBERSHIP_CASES)
def test_membership_metric_restatement_vs_reference_class(name):
    """oracle.membership_losses vs the losses the reference's MembershipLoss returned for the stored sampled images
    and noise (metrics/class_membership.py:69-128): same torch-CPU ops in the same order -> bit-exact."""
    c = load_membership_golden(name)
    got = O.membership_losses(c["all_images"], c["deletion_images"], c["noise"], c["alphas_cumprod"],
                              MembershipStubUNet(), c["timesteps"], c["eval_bs"])
    assert len(got) == len(c["timesteps"]) == c["losses_f32"].shape[0] and len(MEMBERSHIP_CASES) >= 2
    for (a, d), want in zip(got, c["losses_f32"]):
        assert a.dtype == torch.float32 and a.dim() == 0
        assert a.item() == want[0].item() and d.item() == want[1].item()
