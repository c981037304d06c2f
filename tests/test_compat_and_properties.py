"""CPU: the import-path shim (the reference's import line works unchanged), the drop-in class's surface
(method names / parameter names are part of the contract: delete_sd.py calls by keyword), and
hypothesis property tests of the host-side sharding / scaling logic."""
import inspect
import subprocess
import sys
from pathlib import Path

import numpy as np
import torch
from hypothesis import given, settings, strategies as st

ROOT = Path(__file__).resolve().parent.parent


def test_reference_import_line_resolves_to_this_implementation():
    code = ("import sys; sys.path[:0] = [r'%s', r'%s'];"
            "from losses.ddpm_deletion_loss import DDPMDeletionLoss;"
            "import siss_b200.losses.ddpm_deletion_loss as m;"
            "assert DDPMDeletionLoss is m.DDPMDeletionLoss; print('ok')") % (ROOT / "compat", ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", out.stderr


def test_method_surface_matches_reference_contract():
    """Names and parameter order of SURVEY.md §8b (positional call delete_celeb.py:622, keyword call
    delete_sd.py:977-985, **loss_params splat of lambd / superfactor)."""
    from siss_b200.losses import DDPMDeletionLoss
    common = ["self", "unet", "timesteps", "noise", "conditioning", "all_samples_dict", "deletion_samples_dict"]
    want = {"importance_sampling_with_mixture": common + ["lambd"], "double_forward_with_neg_del": common,
            "erasediff": common, "simple_neg_del": common + ["superfactor"], "naive_del": common,
            "subscore_bernoulli": common + ["lambd"]}
    for name, params in want.items():
        got = list(inspect.signature(getattr(DDPMDeletionLoss, name)).parameters)
        assert got[:len(params)] == params, (name, got)
        extra = got[len(params):]
        assert all(inspect.signature(getattr(DDPMDeletionLoss, name)).parameters[e].default is not inspect._empty
                   for e in extra), f"{name}: extra parameters must be optional"
    assert list(inspect.signature(DDPMDeletionLoss.__init__).parameters) == ["self", "gamma", "sigma"]
    obj = DDPMDeletionLoss(gamma=torch.ones(3), sigma=torch.ones(3))
    assert obj.all_gamma is not None and obj.all_sigma is not None          # attribute names of the reference
    assert callable(getattr(obj, "importance_sampling_with_mixture"))        # selected by getattr(..., cfg.deletion.loss_fn)


@settings(max_examples=200, deadline=None)
@given(st.integers(1, 4096), st.integers(1, 16))
def test_shard_bounds_partition(global_batch, world):
    from siss_b200.parallel import shard_bounds
    spans = [shard_bounds(global_batch, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == global_batch
    sizes = [hi - lo for lo, hi in spans]
    assert all(s >= 0 for s in sizes) and max(sizes) - min(sizes) <= 1
    assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


@settings(max_examples=100, deadline=None)
@given(st.integers(1, 512), st.integers(1, 64))
def test_upstream_scale_is_autograds_scalar(train_batch_size, grad_accum):
    """go = d((w.sum()/B)/G)/dw as autograd forms it in fp32."""
    from siss_b200.step import upstream_scale
    w = torch.ones(3, dtype=torch.float32, requires_grad=True)
    ((w.sum() / train_batch_size) / grad_accum).backward()
    assert np.float32(upstream_scale(train_batch_size, grad_accum)) == np.float32(w.grad[0].item())


@settings(max_examples=50, deadline=None)
@given(st.lists(st.integers(1, 5000), min_size=1, max_size=12), st.integers(1, 8))
def test_flat_buffer_layout(sizes, world):
    """GradCombiner's layout: every parameter starts 16-byte aligned, views do not overlap, the total is a
    multiple of 4*world (every rank's shard stays 16-byte aligned)."""
    from siss_b200.grad_combine import GradCombiner
    params = [torch.nn.Parameter(torch.zeros(n)) for n in sizes]
    comb = GradCombiner(params, distributed=False)
    assert comb.total % 4 == 0 and comb.total >= sum(sizes)
    ends = 0
    for off, p in zip(comb.offsets, comb.params):
        assert off % 4 == 0 and off >= ends
        ends = off + p.numel()
    comb.begin_x()
    for p, v in zip(comb.params, comb._views_x):
        assert p.grad.data_ptr() == v.data_ptr() and p.grad.shape == p.shape
    comb.begin_a()
    assert comb.params[0].grad.data_ptr() == comb.g_a.data_ptr()
    quantum = 4 * world
    padded = (comb.total + quantum - 1) // quantum * quantum
    assert padded % quantum == 0
