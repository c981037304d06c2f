"""GPU: the two bindings of the C ABI — the thin torch extension (`torch.ops.siss_b200.*`, csrc/torch_ext.cpp) and the
ctypes table (siss_b200/_lib.py) — call the same entry points and must return bit-identical tensors; the extension is
the default for the hot-path ops (BASELINE.json north_star: "through a thin C-ABI torch extension")."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _both(monkeypatch, fn):
    from siss_b200 import ops
    monkeypatch.setattr(ops, "BINDING", "torch")
    a = fn()
    monkeypatch.setattr(ops, "BINDING", "ctypes")
    b = fn()
    torch.cuda.synchronize()
    a = a if isinstance(a, (tuple, list)) else (a,)
    b = b if isinstance(b, (tuple, list)) else (b,)
    assert len(a) == len(b)
    for i, (u, v) in enumerate(zip(a, b)):
        assert u.dtype == v.dtype and u.shape == v.shape, i
        assert torch.equal(u.view(torch.uint8) if u.dtype.is_floating_point else u,
                           v.view(torch.uint8) if v.dtype.is_floating_point else v), f"output {i} differs between bindings"
    return a


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_hot_ops_identical_through_both_bindings(cuda_device, monkeypatch, dtype):
    from siss_b200 import _lib, ops
    from siss_b200.scheduler import SissDDPMScheduler
    assert ops.BINDING == "torch", "the torch extension is the default binding of the hot ops"
    assert hasattr(_lib.load_ext(), "add_noise_mixture")
    dev = cuda_device
    g = torch.Generator(device=dev).manual_seed(3)
    B, shape = 6, (6, 3, 32, 32)
    x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dtype)
    a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dtype)
    nz = torch.randn(shape, device=dev, generator=g).to(dtype)
    pred, pred2 = torch.randn(shape, device=dev, generator=g), torch.randn(shape, device=dev, generator=g)
    t = torch.randint(0, 1000, (B,), device=dev, generator=g)
    keep = torch.rand(B, generator=torch.Generator().manual_seed(1)) > 0.5           # CPU bool, as the reference draws it
    sched = SissDDPMScheduler()
    ac = sched.alphas_cumprod.to(dev)
    gamma, sigma = sched.gamma_sigma(dev)
    _both(monkeypatch, lambda: ops.add_noise(x0, nz, t, ac))
    xt_x, xt_a = _both(monkeypatch, lambda: ops.add_noise_pair(x0, a0, nz, t, ac))
    _both(monkeypatch, lambda: ops.mixture_weights(xt_x, xt_a, x0, a0, keep, t, gamma, sigma, 0.5))
    x_mix, d_x, d_a, w_x, w_a = _both(monkeypatch, lambda: ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5))
    _both(monkeypatch, lambda: ops.wmse_fwd_bwd(pred, x_mix, x0, a0, t, gamma, sigma, w_x, w_a, 0.125, 0.25))
    g_x, g_a, rl_x, rl_a = _both(monkeypatch, lambda: ops.dual_mse_fwd_bwd(pred, pred2, nz, nz, 0.125, 0.25))
    _both(monkeypatch, lambda: ops.batch_stats(rl_x, rl_a, w_x, w_a, x0[0].numel()))
    _both(monkeypatch, lambda: ops.batch_stats(rl_x, None, None, None, x0[0].numel()))
    G_x, G_a = torch.randn(100_003, device=dev, generator=g), torch.randn(100_003, device=dev, generator=g)
    (sums,) = _both(monkeypatch, lambda: ops.norm3(G_x, G_a))
    _both(monkeypatch, lambda: ops.combine(G_x, G_a, sums, _lib.SISS_COMBINE_SCALING_NORM, 5.0, 1.0))
    _both(monkeypatch, lambda: ops.combine(G_x, G_a, sums, _lib.SISS_COMBINE_ERASEDIFF, 0.1, 1.0, True))


def test_extension_rejects_bad_arguments(cuda_device):
    from siss_b200 import ops
    from siss_b200._lib import SissLibraryError
    dev = cuda_device
    x = torch.zeros(2, 1, 4, 4, device=dev)
    t = torch.zeros(2, dtype=torch.long, device=dev)
    ac = torch.linspace(0.9, 0.1, 10, device=dev)
    with pytest.raises(SissLibraryError):
        ops.add_noise(x.cpu(), x.cpu(), t.cpu(), ac.cpu())          # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        ops.add_noise(x, x[:1], t, ac)                               # shape mismatch
    with pytest.raises(RuntimeError):
        ops.add_noise(x.double(), x.double(), t, ac)                 # unsupported dtype
