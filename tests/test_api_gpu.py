"""GPU: the reference-facing Python surface (DDPMDeletionLoss, GradCombiner, UnlearnStep) driven the
way the task loops drive it (delete_celeb.py:580-767), compared with the golden fixtures and with the
CPU oracle's literal restatement of that loop."""
import copy

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


def _to(d, dev):
    return {k: v.to(dev) for k, v in d.items()}


def _close_w(got, exp, d_x, d_a):
    got, exp = got.cpu().double(), exp.double()
    tol = weight_tolerance(d_x.cpu(), d_a.cpu())
    sat = exp == 0
    assert torch.equal(got[sat], exp[sat])
    assert (((got - exp).abs() / exp.clamp_min(1e-300))[~sat] <= tol[~sat]).all()


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_drop_in_siss_matches_reference_outputs(name, dev):
    """Same seeds, same call, same 7-tuple as the reference run that produced the fixture."""
    from siss_b200.losses import DDPMDeletionLoss
    c = load_golden(name)
    loss = DDPMDeletionLoss(gamma=c["gamma"].to(dev), sigma=c["sigma"].to(dev))
    unet = O.StubUNet().to(dev)
    all_d = _to({"og_latents": c["x0"], "noisy_latents": c["xt_x"]}, dev)
    del_d = _to({"og_latents": c["a0"], "noisy_latents": c["xt_a"]}, dev)
    cond = _to(conditioning_of(c), dev)
    torch.manual_seed(c["siss_mask_seed"])      # the Bernoulli mask comes from the CPU generator, as in the reference
    loss_fn = getattr(loss, "importance_sampling_with_mixture")   # selected by name (delete_celeb.py:373-374)
    items = loss_fn(unet, c["t"].to(dev), c["noise"].to(dev), cond, all_d, del_d, lambd=c["lambd"])
    assert len(items) == 7 and items[0] is None
    _, loss_x, loss_a, w_x, w_a, wl_x, wl_a = items
    assert loss_x.dtype == torch.float32 and loss_x.shape == c["x0"].shape and wl_x.requires_grad
    # weights: tolerance from the conditioning of the reference's own formula
    from siss_b200 import ops
    _, d_x, d_a, _, _ = ops.mixture_weights(all_d["noisy_latents"], del_d["noisy_latents"], all_d["og_latents"],
                                            del_d["og_latents"], c["siss_keep_mask"], c["t"].to(dev), loss.all_gamma,
                                            loss.all_sigma, c["lambd"])
    _close_w(w_x, c["siss_w_x"], d_x, d_a); _close_w(w_a, c["siss_w_a"], d_x, d_a)
    # element-wise losses do not depend on the weights: bit-exact
    assert torch.equal(loss_x.cpu(), c["siss_loss_x"]) and torch.equal(loss_a.cpu(), c["siss_loss_a"])
    tol = weight_tolerance(d_x.cpu(), d_a.cpu()).max().item()
    torch.testing.assert_close(wl_x.detach().cpu(), c["siss_wl_x"], rtol=tol, atol=1e-30)
    torch.testing.assert_close(wl_a.detach().cpu(), c["siss_wl_a"], rtol=tol, atol=1e-30)

    # the task loop's two backward passes (delete_celeb.py:686-702)
    B = c["x0"].shape[0]
    (wl_x.sum() / B).backward(retain_graph=True)
    gx = [p.grad.clone() for p in unet.parameters()]
    (wl_a.sum() / B).backward()
    ga = [p.grad - g for p, g in zip(unet.parameters(), gx)]
    # oracle with the same mask
    ounet = O.StubUNet()
    oit = O.OracleDeletionLoss(c["gamma"], c["sigma"]).importance_sampling_with_mixture(
        ounet, c["t"], c["noise"], conditioning_of(c), {"og_latents": c["x0"], "noisy_latents": c["xt_x"]},
        {"og_latents": c["a0"], "noisy_latents": c["xt_a"]}, lambd=c["lambd"], keep_mask=c["siss_keep_mask"])
    (oit[5].sum() / B).backward(retain_graph=True)
    ogx = [p.grad.clone() for p in ounet.parameters()]
    (oit[6].sum() / B).backward()
    oga = [p.grad - g for p, g in zip(ounet.parameters(), ogx)]
    for a, b in zip(gx + ga, ogx + oga):
        torch.testing.assert_close(a.cpu(), b, rtol=max(tol, 1e-4), atol=1e-5)


@pytest.mark.parametrize("name", ["tshirt_fp32", "celeb_bf16_t999", "sd_fp32_cond", "fp16_mid_t", "odd_D_fp32"])
def test_drop_in_other_methods(name, dev):
    from siss_b200.losses import DDPMDeletionLoss
    c = load_golden(name)
    loss = DDPMDeletionLoss(gamma=c["gamma"].to(dev), sigma=c["sigma"].to(dev))
    all_d = _to({"og_latents": c["x0"], "noisy_latents": c["xt_x"]}, dev)
    del_d = _to({"og_latents": c["a0"], "noisy_latents": c["xt_a"]}, dev)
    cond = _to(conditioning_of(c), dev)
    t, noise = c["t"].to(dev), c["noise"].to(dev)

    it = loss.double_forward_with_neg_del(O.StubUNet().to(dev), t, noise, cond, all_d, del_d)
    assert it[0] is None and it[3] is None and it[4] is None and it[5] is it[1] and it[6] is it[2]
    assert torch.equal(it[1].cpu(), c["nois_loss_x"]) and torch.equal(it[2].cpu(), c["nois_loss_a"])

    it = loss.simple_neg_del(O.StubUNet().to(dev), t, noise, cond, all_d, del_d, superfactor=1.7)
    assert it[1] is None and all(v is None for v in it[3:])
    assert torch.equal(it[0].cpu(), c["neg_loss"]) and torch.equal(it[2].cpu(), c["neg_loss_a"])

    unet = O.StubUNet().to(dev)
    it = loss.naive_del(unet, t, noise, cond, all_d, del_d)
    assert it[1] is it[0] and all(v is None for v in it[2:])
    assert torch.equal(it[0].cpu(), c["naive_loss"])
    (it[0].sum() / 4).backward()
    ounet = O.StubUNet()
    oit = O.OracleDeletionLoss(c["gamma"], c["sigma"]).naive_del(
        ounet, c["t"], c["noise"], conditioning_of(c), {"og_latents": c["x0"], "noisy_latents": c["xt_x"]},
        {"og_latents": c["a0"], "noisy_latents": c["xt_a"]})
    (oit[0].sum() / 4).backward()
    for p, q in zip(unet.parameters(), ounet.parameters()):
        torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=1e-4, atol=1e-5)

    # EraseDiff draws its uniform target on the device generator (as the reference does when run on a
    # GPU); loss_x is deterministic, loss_a is checked against that same device draw.
    torch.manual_seed(99)
    unet = O.StubUNet().to(dev)
    it = loss.erasediff(unet, t, noise, cond, all_d, del_d)
    assert torch.equal(it[1].cpu(), c["erasediff_loss_x"])
    torch.manual_seed(99)
    pred_a = unet(del_d["noisy_latents"], t, **cond)[0]
    u = torch.rand_like(pred_a)
    assert torch.equal(it[2], (pred_a - u) ** 2)


class TinyNet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(5)
        self.c1 = torch.nn.Conv2d(1, 4, 3, padding=1)
        self.c2 = torch.nn.Conv2d(4, 1, 3, padding=1)
        self.odd = torch.nn.Parameter(torch.zeros(3))   # 3 elements: exercises the 16B padding between views

    def forward(self, x, timesteps, return_dict=False, **kw):
        return (self.c2(torch.tanh(self.c1(x))) + self.odd.sum(),)


@pytest.mark.parametrize("loss_fn,kw", [
    ("importance_sampling_with_mixture", dict(lambd=0.5, scaling_norm=5.0)),
    ("double_forward_with_neg_del", dict(scaling_norm=500.0)),
    ("naive_del", dict()),
    ("simple_neg_del", dict(superfactor=0.3)),
    ("subscore_bernoulli", dict(lambd=0.5, scaling_norm=5.0)),
])
@pytest.mark.parametrize("G", [1, 3])
def test_unlearn_step_matches_reference_loop(loss_fn, kw, G, dev):
    """Fast path (K1oK2, K3, dual buffers, K4) == the reference's loop (clone / subtract / dict
    accumulate / per-tensor norms / clip) on the same batches and the same Bernoulli masks."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    B = 4
    cpu_net = TinyNet()
    gpu_net = copy.deepcopy(cpu_net).to(dev)
    sched = SissDDPMScheduler()
    gamma, sigma = O.gamma_sigma(sched.alphas_cumprod)
    oloss = O.OracleDeletionLoss(gamma, sigma)
    loop = O.ReferenceGradLoop(cpu_net, train_batch_size=B, grad_accum_steps=G)
    comb = GradCombiner(gpu_net.parameters())
    step = UnlearnStep(gpu_net, sched, comb, loss_fn=loss_fn, train_batch_size=B, gradient_accumulation_steps=G,
                       max_norm=1.0, **kw)
    torch.manual_seed(77)
    for k in range(G):
        x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
        noise, t = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,))
        keep = torch.rand(B) > 0.5
        all_d = {"og_latents": x0, "noisy_latents": O.add_noise(sched.alphas_cumprod, x0, noise, t)}
        del_d = {"og_latents": a0, "noisy_latents": O.add_noise(sched.alphas_cumprod, a0, noise, t)}
        okw = {}
        if loss_fn in ("importance_sampling_with_mixture", "subscore_bernoulli"):
            if loss_fn == "subscore_bernoulli" and k == G - 1 and G > 1:
                keep = torch.ones(B, dtype=torch.bool)             # no forget row drawn: the second term is a constant
            okw = dict(lambd=0.5, keep_mask=keep)
        elif loss_fn == "simple_neg_del":
            okw = dict(superfactor=0.3)
        items = getattr(oloss, loss_fn)(cpu_net, t, noise, {}, all_d, del_d, **okw)
        loop.micro_step(items, retain_graph=(loss_fn in ("importance_sampling_with_mixture", "subscore_bernoulli")))
        out = step.micro_step(x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev), keep_mask=keep)
        if items[1] is not None and "row_loss_x" in out:
            torch.testing.assert_close(out["row_loss_x"].cpu(), items[1].detach().sum(dim=[1, 2, 3]), rtol=1e-4, atol=1e-5)
        if loss_fn == "subscore_bernoulli":
            assert torch.equal(out["keep_mask"].cpu(), keep)
    assert step.is_sync_step
    single = loss_fn in ("naive_del", "simple_neg_del")
    ref = loop.sync_step(single, loss_fn, scaling_norm=kw.get("scaling_norm"), max_norm=1.0)
    stats = step.sync_step().cpu()
    for p, q in zip(gpu_net.parameters(), cpu_net.parameters()):
        assert p.grad.data_ptr() >= comb.g_x.data_ptr()          # grads live in the flat buffer
        torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=2e-4, atol=2e-6)
    if not single:
        torch.testing.assert_close(stats[0], ref["norm_x"], rtol=1e-4, atol=0)
        torch.testing.assert_close(stats[1], ref["norm_a"], rtol=1e-4, atol=0)
        torch.testing.assert_close(stats[2], ref["scaling_factor"].float(), rtol=1e-4, atol=0)
    torch.testing.assert_close(stats[3], ref["total_norm"], rtol=1e-4, atol=1e-7)
    # second optimiser step reuses the buffers: G_a was cleared, G_x is cleared on begin_x
    assert comb.g_a.abs().max().item() == 0.0


def test_drop_in_loop_with_grad_combiner(dev):
    """The reference loop written with the drop-in class + GradCombiner (the INTEGRATION.md recipe):
    7-tuple -> .sum()/B -> backward twice with begin_x/begin_a -> combine."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.losses import DDPMDeletionLoss
    from siss_b200.scheduler import SissDDPMScheduler
    torch.backends.cudnn.allow_tf32 = False
    B, G = 4, 2
    cpu_net = TinyNet(); gpu_net = copy.deepcopy(cpu_net).to(dev)
    sched = SissDDPMScheduler()
    gamma, sigma = sched.gamma_sigma(dev)
    loss_fn = DDPMDeletionLoss(gamma=gamma, sigma=sigma).importance_sampling_with_mixture
    oloss = O.OracleDeletionLoss(*O.gamma_sigma(sched.alphas_cumprod))
    loop = O.ReferenceGradLoop(cpu_net, train_batch_size=B, grad_accum_steps=G)
    comb = GradCombiner(gpu_net.parameters())
    torch.manual_seed(5)
    for k in range(G):
        x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
        noise, t = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,))
        keep = torch.rand(B) > 0.5
        xt_x, xt_a = sched.add_noise_pair(x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev))
        items = loss_fn(gpu_net, t.to(dev), noise.to(dev), {}, {"og_latents": x0.to(dev), "noisy_latents": xt_x},
                        {"og_latents": a0.to(dev), "noisy_latents": xt_a}, lambd=0.5, keep_mask=keep)
        comb.begin_x(); (items[5].sum() / B / G).backward(retain_graph=True)
        comb.begin_a(); (items[6].sum() / B / G).backward()
        oit = oloss.importance_sampling_with_mixture(
            cpu_net, t, noise, {}, {"og_latents": x0, "noisy_latents": O.add_noise(sched.alphas_cumprod, x0, noise, t)},
            {"og_latents": a0, "noisy_latents": O.add_noise(sched.alphas_cumprod, a0, noise, t)}, lambd=0.5, keep_mask=keep)
        loop.micro_step(oit, retain_graph=True)
    comb.combine(scaling_norm=5.0, max_norm=1.0)
    loop.sync_step(False, scaling_norm=5.0, max_norm=1.0)
    for p, q in zip(gpu_net.parameters(), cpu_net.parameters()):
        torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=2e-4, atol=2e-6)
    st = comb.stats_dict()
    assert set(st) >= {"gradient/norm_loss_x", "gradient/norm_loss_a", "gradient/scaling_factor"}


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_subscore_bernoulli_matches_reference(name, dev):
    """§8(f)3: the reviewer-proposed loss, against the reference's own outputs (bit-exact: select + sqerr)."""
    from siss_b200.losses import DDPMDeletionLoss
    c = load_golden(name)
    loss = DDPMDeletionLoss(gamma=c["gamma"].to(dev), sigma=c["sigma"].to(dev))
    all_d = _to({"og_latents": c["x0"], "noisy_latents": c["xt_x"]}, dev)
    del_d = _to({"og_latents": c["a0"], "noisy_latents": c["xt_a"]}, dev)
    torch.manual_seed(c["subscore_seed"])
    args = (O.StubUNet().to(dev), c["t"].to(dev), c["noise"].to(dev), _to(conditioning_of(c), dev), all_d, del_d)
    if "subscore_raises" in c:
        with pytest.raises(ZeroDivisionError):
            loss.subscore_bernoulli(*args, lambd=c["lambd"])
        return
    it = loss.subscore_bernoulli(*args, lambd=c["lambd"])
    assert it[0] is None and it[3] is None and it[4] is None and it[5] is it[1] and it[6] is it[2]
    assert it[1].shape == c["subscore_loss_x"].shape and it[2].shape == c["subscore_loss_a"].shape
    assert torch.equal(it[1].detach().cpu(), c["subscore_loss_x"])
    assert torch.equal(it[2].detach().cpu(), c["subscore_loss_a"])
    if it[1].numel() > 1 and it[1].requires_grad:
        (it[5].sum() / 4).backward(retain_graph=True)     # the tshirt loop keeps the graph for this method too
        (it[6].sum() / 4).backward()


@pytest.mark.parametrize("name", ["tshirt_fp32", "celeb_bf16_t999", "celeb_fp32_t999", "fp16_mid_t", "odd_D_fp32"])
def test_fused_batch_stats_match_reference_stats_block(name, dev):
    """§8(f)1: the 16 logging scalars (delete_celeb.py:626-656) from the O(B) row sums vs the same
    statistics computed the reference's way on the full [B,C,H,W] tensors of the golden fixture."""
    from siss_b200 import ops
    from siss_b200.step import StepLog, batch_stats
    c = load_golden(name)
    items = (None, c["siss_loss_x"], c["siss_loss_a"], c["siss_w_x"], c["siss_w_a"], None, None)
    ref = O.batch_stats(items)
    D = c["x0"][0].numel()
    out = {"row_loss_x": c["siss_loss_x"].sum(dim=[1, 2, 3]).to(dev), "row_loss_a": c["siss_loss_a"].sum(dim=[1, 2, 3]).to(dev),
           "w_x": c["siss_w_x"].to(dev), "w_a": c["siss_w_a"].to(dev)}
    got = dict(zip(ops.STAT_KEYS, batch_stats(out, D).tolist()))
    for k, v in ref.items():
        assert got[k] == pytest.approx(v, rel=2e-5, abs=1e-30), k
    # missing inputs -> NaN block; B == 1 -> std NaN like torch.std
    part = ops.batch_stats(out["row_loss_x"][:1].contiguous(), None, None, None, D).tolist()
    assert part[0] == pytest.approx(float(out["row_loss_x"][0]) / D, rel=1e-6) and part[3] != part[3] and part[4] != part[4]
    # sync-free logging record
    log = StepLog(depth=2)
    for step in range(3):
        log.push(batch_stats(out, D), torch.arange(5, dtype=torch.float32, device=dev), step)
    recs = log.pop_ready(wait=True)
    assert [r["step"] for r in recs] == [1, 2] and recs[-1]["gradient/scaling_factor"] == 2.0
    assert recs[-1]["loss_x/mean"] == pytest.approx(ref["loss_x/mean"], rel=2e-5)


@pytest.mark.parametrize("loss_fn,kw", [("importance_sampling_with_mixture", dict(lambd=0.5, scaling_norm=5.0)),
                                        ("naive_del", dict())])
def test_fused_combine_adamw_matches_reference_loop_plus_torch_adamw(loss_fn, kw, dev):
    """§8(f)2: three optimiser steps of [reference loop + clip + torch.optim.AdamW + zero_grad] on CPU vs
    [UnlearnStep.micro_step + FusedCombineAdamW.step] on the GPU. lr is large so the update is well above
    fp32 noise in the parameters."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.optim import FusedCombineAdamW
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    B, G = 4, 2
    hp = dict(lr=3e-3, betas=(0.95, 0.999), eps=1e-8, weight_decay=1e-2)
    cpu_net = TinyNet(); gpu_net = copy.deepcopy(cpu_net).to(dev)
    sched = SissDDPMScheduler()
    oloss = O.OracleDeletionLoss(*O.gamma_sigma(sched.alphas_cumprod))
    loop = O.ReferenceGradLoop(cpu_net, train_batch_size=B, grad_accum_steps=G)
    ref_opt = torch.optim.AdamW(cpu_net.parameters(), **hp)
    comb = GradCombiner(gpu_net.parameters())
    opt = FusedCombineAdamW(comb, **hp)
    step = UnlearnStep(gpu_net, sched, comb, loss_fn=loss_fn, train_batch_size=B, gradient_accumulation_steps=G,
                       max_norm=1.0, **kw)
    single = loss_fn == "naive_del"
    torch.manual_seed(31)
    for it in range(3):
        for k in range(G):
            x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
            noise, t = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,))
            keep = torch.rand(B) > 0.5
            all_d = {"og_latents": x0, "noisy_latents": O.add_noise(sched.alphas_cumprod, x0, noise, t)}
            del_d = {"og_latents": a0, "noisy_latents": O.add_noise(sched.alphas_cumprod, a0, noise, t)}
            okw = dict(lambd=0.5, keep_mask=keep) if not single else {}
            loop.micro_step(getattr(oloss, loss_fn)(cpu_net, t, noise, {}, all_d, del_d, **okw), retain_graph=not single)
            step.micro_step(x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev), keep_mask=keep)
        loop.sync_step(single, loss_fn, scaling_norm=kw.get("scaling_norm"), max_norm=1.0)
        ref_opt.step(); ref_opt.zero_grad()
        step._micro = 0
        opt.step(scaling_norm=kw.get("scaling_norm"), max_norm=1.0, single_term=single)
        assert comb.g_x.abs().max().item() == 0.0 and comb.g_a.abs().max().item() == 0.0   # zero_grad folded in
        for p, q in zip(gpu_net.parameters(), cpu_net.parameters()):
            assert p.data_ptr() >= opt.p_flat.data_ptr()                                    # params live in the flat buffer
            torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=3e-5, atol=3e-6)


@pytest.mark.parametrize("decay_on", ["micro_step", "sync_step"])
def test_superfactor_decay_schedules(decay_on, dev):
    """`deletion.superfactor_decay`: delete_celeb.py / delete_tshirt.py multiply loss_params.superfactor after every
    micro-step's statistics (:658-662), delete_sd.py once per optimiser step (delete_sd.py:1173-1193). Two optimiser steps
    of G = 2 with NegGrad against the reference loop fed the superfactor sequence of the respective task file."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    B, G, s0, d = 4, 2, 0.8, 0.5
    cpu_net = TinyNet(); gpu_net = copy.deepcopy(cpu_net).to(dev)
    sched = SissDDPMScheduler()
    oloss = O.OracleDeletionLoss(*O.gamma_sigma(sched.alphas_cumprod))
    loop = O.ReferenceGradLoop(cpu_net, train_batch_size=B, grad_accum_steps=G)
    comb = GradCombiner(gpu_net.parameters())
    step = UnlearnStep(gpu_net, sched, comb, loss_fn="simple_neg_del", train_batch_size=B, gradient_accumulation_steps=G,
                       superfactor=s0, superfactor_decay=d, superfactor_decay_on=decay_on, max_norm=1.0)
    want_sf = [s0, s0 * d, s0 * d * d, s0 * d ** 3] if decay_on == "micro_step" else [s0, s0, s0 * d, s0 * d]
    torch.manual_seed(4)
    for it in range(2):
        for k in range(G):
            sf = want_sf[it * G + k]
            assert step.superfactor == pytest.approx(sf)
            x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
            noise, t = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,))
            all_d = {"og_latents": x0, "noisy_latents": O.add_noise(sched.alphas_cumprod, x0, noise, t)}
            del_d = {"og_latents": a0, "noisy_latents": O.add_noise(sched.alphas_cumprod, a0, noise, t)}
            loop.micro_step(oloss.simple_neg_del(cpu_net, t, noise, {}, all_d, del_d, superfactor=sf), retain_graph=False)
            step.micro_step(x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev))
        loop.sync_step(True, "simple_neg_del", max_norm=1.0)
        step.sync_step()
        for p, q in zip(gpu_net.parameters(), cpu_net.parameters()):
            torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=2e-4, atol=2e-6)
        for q in cpu_net.parameters():
            q.grad = None
    with pytest.raises(ValueError):
        UnlearnStep(gpu_net, sched, comb, loss_fn="naive_del", train_batch_size=B, superfactor_decay_on="epoch")


def test_grad_combiner_after_spelling_equals_begin_spelling(dev):
    """SURVEY §8b sketches the combine boundary as after_backward_x / after_backward_a / combine(mode, value); both
    spellings must drive the same kernels to bit-identical gradients over 2 optimiser steps x 2 micro-steps."""
    from siss_b200.grad_combine import GradCombiner
    torch.manual_seed(9)
    net_a = TinyNet().to(dev); net_b = copy.deepcopy(net_a)
    ca, cb = GradCombiner(net_a.parameters()), GradCombiner(net_b.parameters())
    for it in range(2):
        for k in range(2):
            x = torch.randn(4, 1, 8, 8, device=dev); y = torch.randn(4, 1, 8, 8, device=dev)
            la = net_a(x, None)[0]; lb = net_b(x, None)[0]
            ca.begin_x(); la.square().sum().backward(retain_graph=True)
            ca.begin_a(); (la - y).square().sum().backward()
            lb.square().sum().backward(retain_graph=True); cb.after_backward_x()
            (lb - y).square().sum().backward(); cb.after_backward_a()
        sa = ca.combine(scaling_norm=5.0, max_norm=1.0) if it == 0 else ca.combine(eta=0.05, max_norm=1.0)
        sb = cb.combine(mode="scaling_norm", value=5.0) if it == 0 else cb.combine(mode="erasediff", value=0.05)
        assert torch.equal(sa, sb) and sa[1].item() > 0
        for p, q in zip(net_a.parameters(), net_b.parameters()):
            assert torch.equal(p.grad, q.grad) and p.grad.abs().sum().item() > 0
        cb.zero_grad()                                     # the begin_* flow clears lazily in begin_x()
        assert all(q.grad.abs().sum().item() == 0 for q in net_b.parameters())
    with pytest.raises(ValueError):
        cb.combine(mode="scaling_norm", value=5.0, eta=0.1)
    with pytest.raises(ValueError):
        cb.combine(mode="plain_neg_del", value=1.0)


@pytest.mark.parametrize("device_schedule", [False, True])
def test_fused_adamw_ema_and_lr_schedule(device_schedule, dev):
    """§8(f)2 "(+EMA)": four optimiser steps with a decaying learning rate (lr_scheduler.step(), delete_celeb.py:770)
    and the EMA shadow update (ema_model.step, :776-777; warm-up keys of train_tshirt_mnist.yaml:94-97) folded into
    the optimiser kernel, vs reference loop + torch.optim.AdamW + LambdaLR + the oracle's EMAModel restatement.
    With device_schedule the kernel reads {lr, ema_decay} from device memory."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.optim import FusedCombineAdamW
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    B = 4
    hp = dict(lr=3e-3, betas=(0.95, 0.999), eps=1e-8, weight_decay=1e-2)
    ema_kw = dict(decay=0.9999, use_ema_warmup=True, inv_gamma=1.0, power=0.75)
    lr_at = lambda it: 1.0 - 0.2 * it
    cpu_net = TinyNet(); gpu_net = copy.deepcopy(cpu_net).to(dev)
    sched = SissDDPMScheduler()
    oloss = O.OracleDeletionLoss(*O.gamma_sigma(sched.alphas_cumprod))
    loop = O.ReferenceGradLoop(cpu_net, train_batch_size=B, grad_accum_steps=1)
    ref_opt = torch.optim.AdamW(cpu_net.parameters(), **hp)
    ref_sched = torch.optim.lr_scheduler.LambdaLR(ref_opt, lr_at)
    ref_ema = O.OracleEMA(cpu_net.parameters(), **ema_kw)
    comb = GradCombiner(gpu_net.parameters())
    opt = FusedCombineAdamW(comb, ema=ema_kw, device_schedule=device_schedule, **hp)
    step = UnlearnStep(gpu_net, sched, comb, loss_fn="importance_sampling_with_mixture", train_batch_size=B,
                       lambd=0.5, scaling_norm=5.0, max_norm=1.0)
    torch.manual_seed(77)
    for it in range(4):
        x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
        noise, t = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,))
        keep = torch.rand(B) > 0.5
        all_d = {"og_latents": x0, "noisy_latents": O.add_noise(sched.alphas_cumprod, x0, noise, t)}
        del_d = {"og_latents": a0, "noisy_latents": O.add_noise(sched.alphas_cumprod, a0, noise, t)}
        loop.micro_step(oloss.importance_sampling_with_mixture(cpu_net, t, noise, {}, all_d, del_d, lambd=0.5,
                                                               keep_mask=keep), retain_graph=True)
        loop.sync_step(False, "importance_sampling_with_mixture", scaling_norm=5.0, max_norm=1.0)
        ref_opt.step(); ref_sched.step(); ref_opt.zero_grad()
        ref_ema.step(cpu_net.parameters())
        step.micro_step(x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev), keep_mask=keep)
        step._micro = 0
        opt.set_schedule(lr=hp["lr"] * lr_at(it))
        opt.step(scaling_norm=5.0, max_norm=1.0)
        assert opt.cur_ema_decay == ref_ema.cur_decay_value
        shadow = [opt.ema_flat[o:o + p.numel()].view_as(p) for p, o in zip(comb.params, comb.offsets)]
        for p, q, e, f in zip(gpu_net.parameters(), cpu_net.parameters(), shadow, ref_ema.shadow_params):
            torch.testing.assert_close(p.detach().cpu(), q.detach(), rtol=3e-5, atol=3e-6)
            torch.testing.assert_close(e.cpu(), f, rtol=3e-5, atol=3e-6)
    assert 0.5 < opt.cur_ema_decay < 0.7                       # 1 - (1+3)^-0.75 = 0.646: warm-up is active
    # EMAModel.store / copy_to / restore around evaluation (delete_celeb.py:380-382)
    before = opt.p_flat.clone()
    opt.ema_copy_to_params(); assert torch.equal(opt.p_flat, opt.ema_flat)
    opt.ema_restore_params(); assert torch.equal(opt.p_flat, before)


def test_ema_shadow_update_is_bitwise_the_eager_expression(dev):
    """The in-kernel EMA is s.sub_((1-decay) * (s - p_new)) in fp32 without contraction: with zero gradients, zero
    weight decay and lr = 0 the parameters stay put and the shadow must equal the eager expression bit for bit."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.optim import FusedCombineAdamW
    torch.manual_seed(5)
    net = torch.nn.Linear(257, 33).to(dev)
    comb = GradCombiner(net.parameters())
    opt = FusedCombineAdamW(comb, lr=0.0, weight_decay=0.0, ema=dict(decay=0.9))
    opt.ema_flat.normal_()
    want = opt.ema_flat.clone()
    for it in range(1, 4):
        opt.step(single_term=True, max_norm=None)
        d = (1 + (it - 1)) / (10 + (it - 1)) if it > 1 else 0.0
        want.sub_((1 - min(d, 0.9)) * (want - opt.p_flat))
        assert torch.equal(opt.ema_flat, want), it


@pytest.mark.parametrize("G", [1, 2])
def test_unlearn_step_erasediff_matches_reference_loop(G, dev):
    """EraseDiff on the fast path (dual-MSE kernel, eta-mode combine: s = -max(eta - <X,A>/||A||^2, 0),
    delete_celeb.py:741-742) vs the reference loop, sharing the uniform forget target explicitly."""
    from siss_b200.grad_combine import GradCombiner
    from siss_b200.scheduler import SissDDPMScheduler
    from siss_b200.step import UnlearnStep
    torch.backends.cudnn.allow_tf32 = False
    B, eta = 4, 0.05
    cpu_net = TinyNet(); gpu_net = copy.deepcopy(cpu_net).to(dev)
    sched = SissDDPMScheduler()
    oloss = O.OracleDeletionLoss(*O.gamma_sigma(sched.alphas_cumprod))
    loop = O.ReferenceGradLoop(cpu_net, train_batch_size=B, grad_accum_steps=G)
    comb = GradCombiner(gpu_net.parameters())
    step = UnlearnStep(gpu_net, sched, comb, loss_fn="erasediff", train_batch_size=B, gradient_accumulation_steps=G,
                       eta=eta, max_norm=1.0)
    torch.manual_seed(3)
    for k in range(G):
        x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
        noise, t, u = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,)), torch.rand(B, 1, 8, 8)
        all_d = {"og_latents": x0, "noisy_latents": O.add_noise(sched.alphas_cumprod, x0, noise, t)}
        del_d = {"og_latents": a0, "noisy_latents": O.add_noise(sched.alphas_cumprod, a0, noise, t)}
        loop.micro_step(oloss.erasediff(cpu_net, t, noise, {}, all_d, del_d, uniform_noise=u), retain_graph=False)
        step.micro_step(x0.to(dev), a0.to(dev), noise.to(dev), t.to(dev), forget_target=u.to(dev))
    ref = loop.sync_step(False, "erasediff", eta=eta, max_norm=1.0)
    stats = step.sync_step().cpu()
    torch.testing.assert_close(stats[2], ref["scaling_factor"].float(), rtol=2e-3, atol=1e-5)
    for p, q in zip(gpu_net.parameters(), cpu_net.parameters()):
        torch.testing.assert_close(p.grad.cpu(), q.grad, rtol=2e-3, atol=2e-6)


def test_device_feeder_orders_and_protects_slots(dev):
    """Double-buffered pinned->device feeding: batches come out in submission order with the right contents
    even though the copy of batch i+1 is in flight while batch i is being read, and a slot is never
    overwritten before the compute that read it has been enqueued past it."""
    from siss_b200.feed import DeviceFeeder
    shape = (8, 3, 64, 64)
    feeder = DeviceFeeder([shape, shape], [torch.bfloat16, torch.bfloat16], dev, depth=2)
    hosts = [((torch.full(shape, float(i)).bfloat16()).pin_memory(), (torch.full(shape, float(-i)).bfloat16()).pin_memory())
             for i in range(7)]
    feeder.submit(hosts[0])
    sums = []
    for i in range(7):
        x, a = feeder.next()
        if i + 1 < 7:
            feeder.submit(hosts[i + 1])
        # a long-ish consumer of the slot on the compute stream
        acc = x.float()
        for _ in range(20):
            acc = acc * 1.0 + 0.0
        sums.append((acc.mean() + a.float().mean() * 1000).reshape(1))
    got = torch.cat(sums).cpu()
    exp = torch.tensor([i - 1000.0 * i for i in range(7)])
    torch.testing.assert_close(got, exp, rtol=0, atol=0)
    with pytest.raises(RuntimeError):
        feeder.next()                                   # nothing submitted
    with pytest.raises(ValueError):
        feeder.submit([torch.zeros(shape).bfloat16(), torch.zeros(shape).bfloat16()])   # not pinned


def test_device_feeder_counts_the_handed_out_slot_as_occupied(dev):
    """submit, submit, next, submit (depth 2): the third submit would reuse slot 0, which the consumer of the batch
    just handed out is still reading — it must be refused, and the data of batch 0 must stay intact."""
    from siss_b200.feed import DeviceFeeder
    shape = (4, 3, 32, 32)
    feeder = DeviceFeeder([shape], [torch.float32], dev, depth=2)
    hosts = [[torch.full(shape, float(i)).pin_memory()] for i in range(3)]
    feeder.submit(hosts[0])
    feeder.submit(hosts[1])
    with pytest.raises(RuntimeError):
        feeder.submit(hosts[2])                         # both slots pending
    (x0,) = feeder.next()
    with pytest.raises(RuntimeError):
        feeder.submit(hosts[2])                         # slot 0 handed out and in use, slot 1 pending
    torch.cuda.synchronize()
    assert float(x0.mean()) == 0.0
    (x1,) = feeder.next()                               # releases slot 0
    feeder.submit(hosts[2])                             # now allowed: lands in slot 0
    (x2,) = feeder.next()
    torch.cuda.synchronize()
    assert float(x1.mean()) == 1.0 and float(x2.mean()) == 2.0


@pytest.mark.parametrize("pred_dtype,noise_dtype", [(torch.bfloat16, torch.float32), (torch.float16, torch.bfloat16),
                                                     (torch.float32, torch.bfloat16)])
def test_dual_mse_follows_eager_type_promotion(pred_dtype, noise_dtype, dev):
    """A 16-bit UNet output against an fp32 (or other 16-bit) target: eager computes (pred - target)^2 in the PROMOTED
    dtype and autograd rounds the gradient back to the prediction dtype (ddpm_deletion_loss.py:62-66). The fast path must
    never round the target down to the prediction's 16 bits (advisor finding of round 1)."""
    from siss_b200.step import _dual_mse
    g = torch.Generator(device=dev).manual_seed(9)
    shape = (3, 2, 8, 8)
    px = torch.randn(shape, device=dev, generator=g).to(pred_dtype).requires_grad_(True)
    pa = torch.randn(shape, device=dev, generator=g).to(pred_dtype).requires_grad_(True)
    noise = torch.randn(shape, device=dev, generator=g).to(noise_dtype)
    go = float(np.float32(1.0) / np.float32(3))
    g_x, g_a, rl_x, rl_a = _dual_mse(px.detach(), pa.detach(), noise, noise, go, go)
    lx, la = (px - noise) ** 2, (pa - noise) ** 2                   # eager: promoted dtype
    (lx.sum() / 3).backward()
    (la.sum() / 3).backward()
    assert g_x.dtype == pred_dtype and g_a.dtype == pred_dtype
    assert torch.equal(g_x, px.grad) and torch.equal(g_a, pa.grad)
    torch.testing.assert_close(rl_x.double(), lx.detach().double().sum(dim=[1, 2, 3]), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("kw", [dict(scaling_norm=5.0), dict(eta=5.0)])
def test_grad_combiner_under_a_grad_scaler(kw, dev):
    """mixed_precision: fp16 (delete_celeb.py:104): a real torch GradScaler around the drop-in loop —
    scaler.scale(loss).backward() twice, combine(loss_scale=scaler.get_scale()), scaler.step(optimizer) — against the oracle's
    restatement of the reference's fp16 sequence (norms / factor on scaled gradients, unscale, clip) + a hand-written SGD
    step. The combined gradient is left scaled; the GradScaler's own unscale inside step() finishes the job."""
    from helpers import GoldenStepNet
    from siss_b200.grad_combine import GradCombiner
    torch.backends.cudnn.allow_tf32 = False
    B, S, lr = 4, 2.0 ** 12, 0.25
    net = GoldenStepNet().to(dev)
    ref = copy.deepcopy(net)
    torch.manual_seed(21)
    x = torch.randn(B, 1, 8, 8, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    e_x, e_a = torch.randn_like(x) * 20, torch.randn_like(x) * 20       # large targets: the clip is active
    # reference side: flat gradients of the scaled losses, then the oracle's fp16 sequence
    flats = []
    for tgt in (e_x, e_a):
        loss = ((ref(x, t)[0] - tgt) ** 2).sum() / B
        gs = torch.autograd.grad(loss * S, list(ref.parameters()))
        flats.append(torch.cat([g.reshape(-1) for g in gs]).cpu())
    exp, nx, na, sf, tn, clip = O.combine_flat(flats[0], flats[1], max_norm=1.0, loss_scale=S, **kw)
    assert float(clip) < 1.0 and float(sf) != 0.0
    # product side
    scaler = torch.amp.GradScaler("cuda", init_scale=S)
    opt = torch.optim.SGD(net.parameters(), lr=lr)
    comb = GradCombiner(net.parameters())
    pred = net(x, t)[0]
    comb.begin_x(); scaler.scale(((pred - e_x) ** 2).sum() / B).backward(retain_graph=True)
    comb.begin_a(); scaler.scale(((pred - e_a) ** 2).sum() / B).backward()
    stats = comb.combine(max_norm=1.0, loss_scale=scaler.get_scale(), **kw).clone()
    before = [p.detach().clone() for p in net.parameters()]
    scaler.step(opt); scaler.update()
    assert scaler.get_scale() == S                                        # no inf found, step not skipped
    got = torch.cat([((b - p.detach()) / lr).reshape(-1) for b, p in zip(before, net.parameters())]).cpu()
    torch.testing.assert_close(got, exp, rtol=2e-4, atol=2e-6)
    torch.testing.assert_close(stats.cpu(), torch.stack([nx, na, sf.float(), tn, clip.float()]), rtol=2e-4, atol=1e-7)
