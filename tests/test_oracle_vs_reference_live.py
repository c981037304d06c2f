"""CPU, build container only: the oracle against the REFERENCE ITSELF on randomized cases.

tests/golden/*.npz pin the oracle on nine fixed cases that travel to the GPU box. Where /root/reference exists (the
container the driver runs the CPU suite in), this module additionally imports the reference's own
``losses/ddpm_deletion_loss.py`` and compares every method — outputs AND autograd gradients into the UNet output —
with the oracle on randomized shapes, dtypes, lambdas and timestep patterns, replaying the reference's RNG draws by
seed. Same torch-CPU ops in the same order -> bit-exact. Skipped when the reference checkout is absent."""
import importlib.util
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import siss_oracle as O

REF = Path("/root/reference/losses/ddpm_deletion_loss.py")
pytestmark = pytest.mark.skipif(not REF.is_file(), reason="reference checkout not present (GPU box)")


def _ref_class():
    spec = importlib.util.spec_from_file_location("ref_ddpm_deletion_loss_live", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.DDPMDeletionLoss


class _Stub(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.scale = torch.nn.Parameter(torch.tensor(0.75))
        self.bias = torch.nn.Parameter(torch.tensor(0.05))
        self.preds = []

    def forward(self, x, timesteps, return_dict=False, **kw):
        out = x.to(torch.float32) * self.scale + self.bias + timesteps.reshape(-1, 1, 1, 1).float() * 1e-4
        out.retain_grad()
        self.preds.append(out)
        return (out,)


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    for i in range(n):
        shape = (int(rng.integers(1, 9)), int(rng.integers(1, 5)), int(rng.integers(1, 20)), int(rng.integers(1, 20)))
        dtype = [torch.float32, torch.bfloat16, torch.float16][int(rng.integers(0, 3))]
        lambd = float(rng.choice([0.0, 0.1, 0.3, 0.5, 0.9, 1.0]))
        tmode = int(rng.integers(0, 4))
        schedule = ["linear", "scaled_linear"][int(rng.integers(0, 2))]
        yield pytest.param(shape, dtype, lambd, tmode, schedule, 1000 + i,
                           id=f"{'x'.join(map(str, shape))}-{str(dtype).split('.')[-1]}-l{lambd}-t{tmode}-{schedule}")


def _eq(a, b, what):
    if a is None or b is None:
        assert a is None and b is None, what
        return
    assert a.dtype == b.dtype and a.shape == b.shape, what
    assert torch.equal(a.float().nan_to_num(nan=-7.0, posinf=9e30, neginf=-9e30),
                       b.float().nan_to_num(nan=-7.0, posinf=9e30, neginf=-9e30)), what


@pytest.mark.parametrize("shape,dtype,lambd,tmode,schedule,seed", list(_cases(36, seed=7)))
def test_oracle_equals_reference_on_random_cases(shape, dtype, lambd, tmode, schedule, seed):
    Ref = _ref_class()
    B = shape[0]
    torch.manual_seed(seed)
    x0 = (torch.rand(shape) * 2 - 1).to(dtype); a0 = (torch.rand(shape) * 2 - 1).to(dtype); noise = torch.randn(shape).to(dtype)
    t = [torch.randint(0, 1000, (B,)), torch.full((B,), 999), torch.randint(0, 30, (B,)), torch.randint(600, 1000, (B,))][tmode]
    kw = dict(beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012) if schedule == "scaled_linear" else {}
    ac = O.make_alphas_cumprod(**kw)
    gamma, sigma = O.gamma_sigma(ac)
    all_d = {"og_latents": x0, "noisy_latents": O.add_noise(ac, x0, noise, t)}
    del_d = {"og_latents": a0, "noisy_latents": O.add_noise(ac, a0, noise, t)}
    ref, ora = Ref(gamma=gamma, sigma=sigma), O.OracleDeletionLoss(gamma, sigma)

    def run(obj, name, draw_seed, **kwargs):
        net = _Stub()
        torch.manual_seed(draw_seed)                         # both sides make the same torch RNG draws in the same order
        items = getattr(obj, name)(net, t, noise, {}, all_d, del_d, **kwargs)
        grads = []
        scalars = [v for v in (items[0], items[5], items[6]) if v is not None and v.numel() > 0]
        for k, sc in enumerate(scalars):
            for p in net.preds:
                p.grad = None
            (sc.sum() / B).backward(retain_graph=k + 1 < len(scalars))
            grads.append([None if p.grad is None else p.grad.clone() for p in net.preds])
        return items, grads

    methods = [("importance_sampling_with_mixture", dict(lambd=lambd)), ("double_forward_with_neg_del", {}),
               ("erasediff", {}), ("simple_neg_del", dict(superfactor=1.7)), ("naive_del", {})]
    if lambd < 1.0:                                          # the reference divides by (1 - lambd) in Python
        methods.append(("subscore_bernoulli", dict(lambd=lambd)))
    for name, kwargs in methods:
        r_items, r_grads = run(ref, name, seed + 17, **kwargs)
        o_items, o_grads = run(ora, name, seed + 17, **kwargs)
        assert len(r_items) == len(o_items) == 7
        for i, (a, b) in enumerate(zip(r_items, o_items)):
            _eq(a, b, f"{name}: item {i}")
        assert len(r_grads) == len(o_grads)
        for gi, (ga, gb) in enumerate(zip(r_grads, o_grads)):
            for pi, (a, b) in enumerate(zip(ga, gb)):
                _eq(a, b, f"{name}: grad of scalar {gi} into pred {pi}")
    if lambd == 1.0:
        for obj in (ref, ora):
            with pytest.raises(ZeroDivisionError):
                obj.subscore_bernoulli(_Stub(), t, noise, {}, all_d, del_d, lambd=1.0)


REF_METRIC = Path("/root/reference/metrics/class_membership.py")


@pytest.mark.skipif(not REF_METRIC.is_file(), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("seed", range(8))
def test_membership_restatement_equals_reference_on_random_cases(seed):
    """oracle.membership_losses vs the reference's MembershipLoss.compute_membership_losses on random grids."""
    spec = importlib.util.spec_from_file_location("ref_class_membership_live", REF_METRIC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(seed)
    shape = (int(rng.integers(1, 4)), int(rng.integers(2, 12)), int(rng.integers(2, 12)))
    n_img, n_noise, eval_bs = int(rng.integers(1, 7)), int(rng.integers(1, 6)), int(rng.integers(1, 9))
    timesteps = [int(v) for v in rng.integers(0, 1000, size=int(rng.integers(1, 4)))]
    torch.manual_seed(seed)
    ac = O.make_alphas_cumprod()

    class Sched:
        alphas_cumprod = ac

        @staticmethod
        def add_noise(x0, noise, t):
            return O.add_noise(ac, x0, noise, t)

    class Net(torch.nn.Module):
        def forward(self, x, timesteps, return_dict=False, **kw):
            return (x * 0.75 + 0.05 + timesteps.reshape(-1, 1, 1, 1).float() * 1e-4,)

    m = mod.MembershipLoss([torch.rand(shape) for _ in range(n_img + 3)], [torch.rand(shape) for _ in range(n_img + 1)],
                           Sched(), Net(), n_img, n_noise, eval_bs, "cpu")
    m.sample_images(); m.sample_noises()
    want = m.compute_membership_losses(timesteps)
    got = O.membership_losses(m.all_sampled_images, m.deletion_sampled_images, m.noise, ac, Net(), timesteps, eval_bs)
    assert len(got) == len(want) == len(timesteps)
    for (a, d), (ra, rd) in zip(got, want):
        assert a.item() == ra.item() and d.item() == rd.item()
