"""GPU parity for the membership-loss metric (reference metrics/class_membership.py):
the two kernels through the C ABI vs the oracle, and the mirrored MembershipLoss class vs the losses the
reference class itself returned (tests/golden/membership_*.npz).

Tolerances: noisy batches BIT-EXACT (same rounding sequence as add_noise); per-row sums and the final means
rtol 1e-5 (fp32 reduction order)."""
import random

import pytest
import torch

from helpers import MEMBERSHIP_CASES, MembershipStubUNet, load_membership_golden
from oracle import siss_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev(cuda_device):
    from siss_b200 import _lib
    _lib.load()
    return cuda_device


def _expanded(images, noise):
    n_img, n_noise = images.shape[0], noise.shape[0]
    img = images.unsqueeze(1).expand(-1, n_noise, *images.shape[1:]).reshape(-1, *images.shape[1:])
    nz = noise.unsqueeze(0).expand(n_img, *noise.shape).reshape(-1, *noise.shape[1:])
    return img, nz


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape,n_img,n_noise", [((1, 28, 28), 6, 5), ((3, 10, 9), 5, 3), ((3, 64, 64), 3, 4),
                                                  ((1, 7, 3), 9, 1), ((3, 256, 256), 2, 3)])
def test_membership_add_noise_bit_exact(dtype, shape, n_img, n_noise, dev):
    from siss_b200 import ops
    torch.manual_seed(hash((shape, n_img)) % 1000)
    ac = O.make_alphas_cumprod()
    x0 = (torch.rand(n_img, *shape) * 2 - 1).to(dtype)
    a0 = (torch.rand(n_img, *shape) * 2 - 1).to(dtype)
    noise = torch.randn(n_noise, *shape).to(dtype)
    total = n_img * n_noise
    for t in (0, 437, 999):
        xe, nz = _expanded(x0, noise)
        ae, _ = _expanded(a0, noise)
        tt = torch.full((total,), t)
        want_x, want_a = O.add_noise(ac, xe, nz, tt), O.add_noise(ac, ae, nz, tt)
        # whole grid in one call, and ragged slices the way eval batches cut it (class_membership.py:101-105)
        for r0, rows in [(0, total), (1, total - 1), (total // 2, total - total // 2), (total - 1, 1), (0, 0)]:
            gx, ga = ops.membership_add_noise(x0.to(dev), a0.to(dev), noise.to(dev), t, ac, r0, rows)
            assert gx.dtype == dtype and gx.shape == (rows, *shape)
            assert torch.equal(gx.cpu(), want_x[r0:r0 + rows]), (t, r0, rows)
            assert torch.equal(ga.cpu(), want_a[r0:r0 + rows]), (t, r0, rows)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape,rows,n_noise", [((1, 28, 28), 30, 5), ((3, 10, 9), 7, 3), ((3, 256, 256), 5, 2),
                                                 ((3, 512, 512), 2, 2), ((1, 1, 1), 3, 2)])
def test_membership_sqerr_vs_oracle(dtype, shape, rows, n_noise, dev):
    """Rows shorter than, equal to and much longer than a CTA span (row split over many CTAs)."""
    from siss_b200 import ops
    torch.manual_seed(rows)
    noise = torch.randn(n_noise, *shape).to(dtype)
    px, pa = torch.randn(rows, *shape), torch.randn(rows, *shape) * 0.5 + 0.1
    for r0 in (0, 3):
        nz = noise[(torch.arange(rows) + r0) % n_noise].double()
        want_x = ((px.double() - nz) ** 2).flatten(1).sum(1)
        want_a = ((pa.double() - nz) ** 2).flatten(1).sum(1)
        sx, sa = ops.membership_sqerr(px.to(dev), pa.to(dev), noise.to(dev), r0)
        assert sx.dtype == torch.float32 and sx.shape == (rows,)
        torch.testing.assert_close(sx.cpu().double(), want_x, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(sa.cpu().double(), want_a, rtol=1e-5, atol=1e-6)
        # launch-to-launch reproducibility (fixed-order combine of the row partials)
        sx2, sa2 = ops.membership_sqerr(px.to(dev), pa.to(dev), noise.to(dev), r0)
        assert torch.equal(sx, sx2) and torch.equal(sa, sa2)


@pytest.mark.parametrize("name", MEMBERSHIP_CASES)
def test_membership_class_vs_reference_class_outputs(name, dev):
    """Same constructor / methods / return structure as the reference class; losses vs what the reference returned."""
    from siss_b200.metrics import MembershipLoss
    from siss_b200.scheduler import SissDDPMScheduler
    c = load_membership_golden(name)
    ds_all, ds_del = list(c["dataset_all"]), list(c["dataset_deletion"])
    n_img, n_noise = c["all_images"].shape[0], c["noise"].shape[0]
    m = MembershipLoss(ds_all, ds_del, SissDDPMScheduler(), MembershipStubUNet().to(dev), n_img, n_noise, c["eval_bs"], dev)
    random.seed(c["seed"])
    m.sample_images()                                   # same Python-RNG draws as the reference (:30-62)
    assert torch.equal(m.all_sampled_images.cpu(), c["all_images"])
    assert torch.equal(m.deletion_sampled_images.cpu(), c["deletion_images"])
    m.sample_noises()
    assert m.noise.shape == c["noise"].shape and m.noise.device.type == "cuda"
    m.noise = c["noise"].to(dev)                        # device RNG differs from the CPU draw: share the stored noise
    got = m.compute_membership_losses(c["timesteps"])
    assert len(got) == len(c["timesteps"])
    for (a, d), want in zip(got, c["losses"]):
        assert a.dim() == 0 and a.dtype == torch.float32 and a.device.type == "cuda"
        torch.testing.assert_close(a.cpu().double(), want[0], rtol=1e-5, atol=0)
        torch.testing.assert_close(d.cpu().double(), want[1], rtol=1e-5, atol=0)


def test_membership_argument_errors(dev):
    from siss_b200 import ops
    ac = O.make_alphas_cumprod()
    x = torch.zeros(2, 1, 4, 4, device=dev)
    nz = torch.zeros(3, 1, 4, 4, device=dev)
    with pytest.raises(ValueError):
        ops.membership_add_noise(x, x, nz, 5, ac, 4, 3)            # rows past the 2 x 3 grid
    with pytest.raises(IndexError):
        ops.membership_add_noise(x, x, nz, 1000, ac, 0, 6)         # timestep outside the schedule
    with pytest.raises(ValueError):
        ops.membership_add_noise(x, x, nz.bfloat16(), 5, ac, 0, 6)
    with pytest.raises(ValueError):
        ops.membership_sqerr(torch.zeros(2, 1, 4, 5, device=dev), torch.zeros(2, 1, 4, 5, device=dev), nz, 0)
