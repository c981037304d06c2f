"""Tensor-level wrappers over the C ABI (include/siss_b200.h).

PyTorch is used here for device memory, streams and dtype bookkeeping only; every arithmetic
operation on ``[B, D]`` data happens inside libsiss_b200.so. All calls are asynchronous on the
current CUDA stream and never synchronise. CPU tensors are rejected: there is no fallback.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import SISS_BF16, SISS_F16, SISS_F32, SissLibraryError

_DTYPES = {torch.float32: SISS_F32, torch.bfloat16: SISS_BF16, torch.float16: SISS_F16}

# incremented on every kernel launch made through this module (bench.py reports it)
launch_count = 0

# Two bindings of the SAME C ABI (include/siss_b200.h): "torch" = the thin torch extension csrc/torch_ext.cpp
# (`torch.ops.siss_b200.*`: validation, output allocation and pointer marshalling in C++, one dispatcher call per op —
# what BASELINE.json's north_star names), used for the ops on the per-step hot path; "ctypes" = the table in _lib.py,
# the documented second binding and the only one for the cold ops. SISS_BINDING=ctypes forces it
# everywhere (tests run both and compare bit for bit). Neither is a fallback for the other's failure.
BINDING = os.environ.get("SISS_BINDING", "torch").lower()
if BINDING not in ("torch", "ctypes"):
    raise SissLibraryError(f"SISS_BINDING={BINDING!r}: expected 'torch' or 'ctypes'")


def _ext_for(*tensors: torch.Tensor):
    """torch.ops.siss_b200 after the cheap per-call checks the C++ side cannot turn into SissLibraryError."""
    for t in tensors:
        if not t.is_cuda:
            raise SissLibraryError(
                "siss_b200 ops need CUDA tensors on a B200; there is no CPU path (the CPU restatement is "
                "oracle/, for tests only)")
    idx = tensors[0].device.index
    if idx not in _lib._devices_checked:
        _need_cuda(*tensors)
    return _lib.load_ext()


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


def _dt(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise SissLibraryError(f"unsupported dtype {t.dtype}; siss_b200 handles float32/bfloat16/float16") from None


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise SissLibraryError(
                "siss_b200 ops need CUDA tensors on a B200; there is no CPU path (the CPU restatement is "
                "oracle/, for tests only)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise SissLibraryError(f"tensors on different devices: {dev} vs {t.device}")
    if dev is not None:
        if dev.index != torch.cuda.current_device():
            raise SissLibraryError(f"tensors live on {dev} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                                   "kernels launch on the current device's stream (use torch.cuda.device(...))")
        _lib.require_b200(dev.index)
    return dev


def _rows(t: torch.Tensor) -> Tuple[int, int]:
    B = t.shape[0]
    return B, (t.numel() // B if B else 0)


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _timesteps(ts: torch.Tensor, B: int) -> torch.Tensor:
    if ts.dtype != torch.int64:
        ts = ts.long()
    if ts.dim() == 0:
        ts = ts.expand(B)
    return _c(ts)


def _table(tab: torch.Tensor, dev: torch.device) -> torch.Tensor:
    if tab.device != dev or tab.dtype != torch.float32 or not tab.is_contiguous():
        tab = tab.to(device=dev, dtype=torch.float32).contiguous()
    return tab


# ------------------------------------------------------------------------------------------------
# workspaces: one per (device, stream); zeroed once, the kernels leave them clean
# ------------------------------------------------------------------------------------------------
_row_ws = {}
_norm_ws = {}


def _row_workspace(dev: torch.device, B: int) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _row_ws.get(key)
    need = _lib.load().siss_row_workspace_bytes(max(B, 1))
    if ws is None or ws[0] < B:
        cap = max(B, 64)
        nbytes = _lib.load().siss_row_workspace_bytes(cap)
        assert nbytes >= need
        ws = (cap, torch.zeros(nbytes, dtype=torch.uint8, device=dev))
        _row_ws[key] = ws
    return ws[1]


def _norm_workspace(dev: torch.device) -> torch.Tensor:
    key = (dev.index, torch.cuda.current_stream().cuda_stream)
    ws = _norm_ws.get(key)
    if ws is None:
        ws = torch.zeros(_lib.load().siss_norm3_workspace_bytes(), dtype=torch.uint8, device=dev)
        _norm_ws[key] = ws
    return ws


# ------------------------------------------------------------------------------------------------
# K1
# ------------------------------------------------------------------------------------------------
def add_noise(x0: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor,
              alphas_cumprod: torch.Tensor) -> torch.Tensor:
    """x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) eps, rounding like diffusers' DDPMScheduler.add_noise."""
    if BINDING == "torch":
        out = _ext_for(x0, noise).add_noise(x0, noise, timesteps, alphas_cumprod)
        if out.numel():
            _count()
        return out
    dev = _need_cuda(x0, noise, timesteps)
    if noise.shape != x0.shape or noise.dtype != x0.dtype:
        raise ValueError("noise must have the shape and dtype of the samples")
    x0, noise = _c(x0), _c(noise)
    B, D = _rows(x0)
    ts = _timesteps(timesteps, B)
    ac = _table(alphas_cumprod, dev)
    out = torch.empty_like(x0)
    if out.numel() == 0:
        return out
    _lib.check(_lib.load().siss_add_noise(_ptr(x0), _ptr(noise), _ptr(ts), _ptr(ac), ac.numel(), _ptr(out),
                                          B, D, _dt(x0), _stream()), "siss_add_noise")
    _count()
    return out


def add_noise_pair(x0: torch.Tensor, a0: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor,
                   alphas_cumprod: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Keep and forget batches noised with the shared eps and t in one pass (5 streams instead of 6)."""
    if BINDING == "torch":
        out = _ext_for(x0, a0, noise).add_noise_pair(x0, a0, noise, timesteps, alphas_cumprod)
        if out[0].numel():
            _count()
        return out
    dev = _need_cuda(x0, a0, noise, timesteps)
    if not (x0.shape == a0.shape == noise.shape) or not (x0.dtype == a0.dtype == noise.dtype):
        raise ValueError("x0, a0 and noise must share shape and dtype")
    x0, a0, noise = _c(x0), _c(a0), _c(noise)
    B, D = _rows(x0)
    ts = _timesteps(timesteps, B)
    ac = _table(alphas_cumprod, dev)
    xt_x, xt_a = torch.empty_like(x0), torch.empty_like(a0)
    if xt_x.numel() == 0:
        return xt_x, xt_a
    _lib.check(_lib.load().siss_add_noise_pair(_ptr(x0), _ptr(a0), _ptr(noise), _ptr(ts), _ptr(ac), ac.numel(),
                                               _ptr(xt_x), _ptr(xt_a), B, D, _dt(x0), _stream()),
               "siss_add_noise_pair")
    _count()
    return xt_x, xt_a


# ------------------------------------------------------------------------------------------------
# K2
# ------------------------------------------------------------------------------------------------
def _keep_mask(keep_mask: torch.Tensor, dev: torch.device, B: int) -> torch.Tensor:
    if keep_mask.shape != (B,):
        raise ValueError(f"keep_mask must have shape ({B},)")
    if keep_mask.dtype == torch.bool:
        keep_mask = keep_mask.to(torch.uint8)
    elif keep_mask.dtype != torch.uint8:
        raise ValueError("keep_mask must be bool or uint8")
    if keep_mask.device != dev:
        # the reference draws the mask on the CPU (losses/ddpm_deletion_loss.py:18); B bytes H2D
        keep_mask = keep_mask.to(dev, non_blocking=True)
    return _c(keep_mask)


def mixture_weights(xt_x: torch.Tensor, xt_a: torch.Tensor, x0: torch.Tensor, a0: torch.Tensor,
                    keep_mask: torch.Tensor, timesteps: torch.Tensor, gamma: torch.Tensor,
                    sigma: torch.Tensor, lambd: float):
    """Row-select the mixture sample and compute (dist_x, dist_a, w_x, w_a). Returns
    (x_mix, dist_x, dist_a, w_x, w_a)."""
    if BINDING == "torch":
        out = _ext_for(xt_x, xt_a, x0, a0).mixture_weights(xt_x, xt_a, x0, a0, keep_mask, timesteps, gamma, sigma, float(lambd))
        _count()
        return out
    dev = _need_cuda(xt_x, xt_a, x0, a0, timesteps)
    if not (xt_x.shape == xt_a.shape == x0.shape == a0.shape):
        raise ValueError("noisy/original keep/forget batches must share one shape")
    if not (xt_x.dtype == xt_a.dtype == x0.dtype == a0.dtype):
        raise ValueError("noisy/original keep/forget batches must share one dtype")
    xt_x, xt_a, x0, a0 = _c(xt_x), _c(xt_a), _c(x0), _c(a0)
    B, D = _rows(x0)
    ts = _timesteps(timesteps, B)
    keep = _keep_mask(keep_mask, dev, B)
    g, s = _table(gamma, dev), _table(sigma, dev)
    x_mix = torch.empty_like(xt_x)
    small = torch.empty((4, B), dtype=torch.float32, device=dev)
    ws = _row_workspace(dev, B)
    _lib.check(_lib.load().siss_mixture_weights(
        _ptr(xt_x), _ptr(xt_a), _ptr(x0), _ptr(a0), _ptr(keep), _ptr(ts), _ptr(g), _ptr(s), g.numel(),
        float(lambd), _ptr(x_mix), _ptr(small[0]), _ptr(small[1]), _ptr(small[2]), _ptr(small[3]), _ptr(ws),
        B, D, _dt(x0), _stream()), "siss_mixture_weights")
    _count()
    return x_mix, small[0], small[1], small[2], small[3]


def add_noise_mixture(x0: torch.Tensor, a0: torch.Tensor, noise: torch.Tensor, keep_mask: torch.Tensor,
                      timesteps: torch.Tensor, alphas_cumprod: torch.Tensor, gamma: torch.Tensor,
                      sigma: torch.Tensor, lambd: float):
    """Fused K1 o K2: returns (x_mix, dist_x, dist_a, w_x, w_a) straight from (x0, a0, eps)."""
    if BINDING == "torch":
        out = _ext_for(x0, a0, noise).add_noise_mixture(x0, a0, noise, keep_mask, timesteps, alphas_cumprod, gamma, sigma,
                                                        float(lambd))
        _count()
        return out
    dev = _need_cuda(x0, a0, noise, timesteps)
    if not (x0.shape == a0.shape == noise.shape) or not (x0.dtype == a0.dtype == noise.dtype):
        raise ValueError("x0, a0 and noise must share shape and dtype")
    x0, a0, noise = _c(x0), _c(a0), _c(noise)
    B, D = _rows(x0)
    ts = _timesteps(timesteps, B)
    keep = _keep_mask(keep_mask, dev, B)
    ac, g, s = _table(alphas_cumprod, dev), _table(gamma, dev), _table(sigma, dev)
    x_mix = torch.empty_like(x0)
    small = torch.empty((4, B), dtype=torch.float32, device=dev)
    ws = _row_workspace(dev, B)
    _lib.check(_lib.load().siss_add_noise_mixture(
        _ptr(x0), _ptr(a0), _ptr(noise), _ptr(keep), _ptr(ts), _ptr(ac), _ptr(g), _ptr(s), g.numel(),
        float(lambd), _ptr(x_mix), _ptr(small[0]), _ptr(small[1]), _ptr(small[2]), _ptr(small[3]), _ptr(ws),
        B, D, _dt(x0), _stream()), "siss_add_noise_mixture")
    _count()
    return x_mix, small[0], small[1], small[2], small[3]


def add_noise_mixture_rng(x0: torch.Tensor, a0: torch.Tensor, keep_mask: torch.Tensor, timesteps: torch.Tensor,
                          alphas_cumprod: torch.Tensor, gamma: torch.Tensor, sigma: torch.Tensor, lambd: float,
                          seed: int, draw: int, elem_offset: int = 0, want_noise: bool = False,
                          d_draw: Optional[torch.Tensor] = None):
    """K1 o K2 with eps drawn in-kernel from the counter-based stream (opt-in, siss_b200.rng): returns
    (x_mix, dist_x, dist_a, w_x, w_a, noise-or-None). Bit-identical to ``DeviceRng.randn`` + ``add_noise_mixture``."""
    dev = _need_cuda(x0, a0, timesteps)
    if x0.shape != a0.shape or x0.dtype != a0.dtype:
        raise ValueError("x0 and a0 must share shape and dtype")
    x0, a0 = _c(x0), _c(a0)
    B, D = _rows(x0)
    ts = _timesteps(timesteps, B)
    keep = _keep_mask(keep_mask, dev, B)
    ac, g, s = _table(alphas_cumprod, dev), _table(gamma, dev), _table(sigma, dev)
    x_mix = torch.empty_like(x0)
    noise = torch.empty_like(x0) if want_noise else None
    small = torch.empty((4, B), dtype=torch.float32, device=dev)
    if x_mix.numel() == 0:
        return x_mix, small[0], small[1], small[2], small[3], noise
    ws = _row_workspace(dev, B)
    _lib.check(_lib.load().siss_add_noise_mixture_rng(
        _ptr(x0), _ptr(a0), _ptr(keep), _ptr(ts), _ptr(ac), _ptr(g), _ptr(s), g.numel(), float(lambd),
        int(seed) & 0xFFFFFFFFFFFFFFFF, int(draw), _ptr(d_draw), int(elem_offset), _ptr(x_mix), _ptr(noise), _ptr(small[0]),
        _ptr(small[1]), _ptr(small[2]), _ptr(small[3]), _ptr(ws), B, D, _dt(x0), _stream()), "siss_add_noise_mixture_rng")
    _count()
    return x_mix, small[0], small[1], small[2], small[3], noise


# ------------------------------------------------------------------------------------------------
# K3
# ------------------------------------------------------------------------------------------------
def _k3_common(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a):
    dev = _need_cuda(pred, x_mix, x0, a0, timesteps, w_x, w_a)
    if not (pred.shape == x_mix.shape == x0.shape == a0.shape):
        raise ValueError("pred, x_mix, x0, a0 must share one shape")
    if not (x_mix.dtype == x0.dtype == a0.dtype):
        raise ValueError("x_mix, x0, a0 must share one dtype")
    pred, x_mix, x0, a0 = _c(pred), _c(x_mix), _c(x0), _c(a0)
    B, D = _rows(x0)
    ts = _timesteps(timesteps, B)
    g, s = _table(gamma, dev), _table(sigma, dev)
    w_x, w_a = _c(w_x.float()), _c(w_a.float())
    return dev, pred, x_mix, x0, a0, ts, g, s, w_x, w_a, B, D


def wmse_fwd_bwd(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a, go_x: float, go_a: float):
    """K3 fast path. Returns (grad_x, grad_a, row_loss_x, row_loss_a); grads in pred's dtype."""
    if BINDING == "torch":
        out = _ext_for(pred, x_mix, x0, a0, w_x, w_a).wmse_fwd_bwd(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a,
                                                                   float(go_x), float(go_a))
        _count()
        return out
    dev, pred, x_mix, x0, a0, ts, g, s, w_x, w_a, B, D = _k3_common(pred, x_mix, x0, a0, timesteps, gamma, sigma,
                                                                     w_x, w_a)
    grad_x, grad_a = torch.empty_like(pred), torch.empty_like(pred)
    rows = torch.empty((2, B), dtype=torch.float32, device=dev)
    ws = _row_workspace(dev, B)
    _lib.check(_lib.load().siss_wmse_fwd_bwd(
        _ptr(pred), _dt(pred), _ptr(x_mix), _ptr(x0), _ptr(a0), _dt(x0), _ptr(ts), _ptr(g), _ptr(s), g.numel(),
        _ptr(w_x), _ptr(w_a), float(go_x), float(go_a), _ptr(grad_x), _ptr(grad_a), _ptr(rows[0]), _ptr(rows[1]),
        _ptr(ws), B, D, _stream()), "siss_wmse_fwd_bwd")
    _count()
    return grad_x, grad_a, rows[0], rows[1]


def wmse_fwd(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a):
    """Materialise (loss_x, loss_a, w_x*loss_x, w_a*loss_a) as fp32 tensors shaped like pred."""
    dev, pred, x_mix, x0, a0, ts, g, s, w_x, w_a, B, D = _k3_common(pred, x_mix, x0, a0, timesteps, gamma, sigma,
                                                                     w_x, w_a)
    outs = [torch.empty(pred.shape, dtype=torch.float32, device=dev) for _ in range(4)]
    _lib.check(_lib.load().siss_wmse_fwd(
        _ptr(pred), _dt(pred), _ptr(x_mix), _ptr(x0), _ptr(a0), _dt(x0), _ptr(ts), _ptr(g), _ptr(s), g.numel(),
        _ptr(w_x), _ptr(w_a), _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), B, D, _stream()),
        "siss_wmse_fwd")
    _count()
    return tuple(outs)


def _grad_out(go: Optional[torch.Tensor], shape, dtype=torch.float32):
    """Classify an upstream gradient: (tensor_or_None, stride) with stride 0 for a broadcast scalar."""
    if go is None:
        return None, 0
    if go.dtype != dtype:
        go = go.to(dtype)
    if go.numel() == 1 or all(st == 0 for st in go.stride()):
        # `.sum()` backward: an expanded 0-dim value; read the single element on the device
        return go.as_strided((1,), (1,), go.storage_offset()) if go.numel() != 1 else go.reshape(1), 0
    if tuple(go.shape) != tuple(shape):
        go = go.expand(shape)
    return go.contiguous(), 1


def wmse_bwd(pred, x_mix, x0, a0, timesteps, gamma, sigma, w_x, w_a,
             go_loss_x=None, go_loss_a=None, go_wloss_x=None, go_wloss_a=None):
    """General backward of wmse_fwd into pred (dtype of pred)."""
    dev, pred, x_mix, x0, a0, ts, g, s, w_x, w_a, B, D = _k3_common(pred, x_mix, x0, a0, timesteps, gamma, sigma,
                                                                     w_x, w_a)
    gs = [_grad_out(t, pred.shape) for t in (go_loss_x, go_loss_a, go_wloss_x, go_wloss_a)]
    grad = torch.empty_like(pred)
    if all(t is None for t, _ in gs):
        return grad.zero_()
    _need_cuda(*[t for t, _ in gs])
    _lib.check(_lib.load().siss_wmse_bwd(
        _ptr(pred), _dt(pred), _ptr(x_mix), _ptr(x0), _ptr(a0), _dt(x0), _ptr(ts), _ptr(g), _ptr(s), g.numel(),
        _ptr(w_x), _ptr(w_a), _ptr(gs[0][0]), gs[0][1], _ptr(gs[1][0]), gs[1][1], _ptr(gs[2][0]), gs[2][1],
        _ptr(gs[3][0]), gs[3][1], _ptr(grad), B, D, _stream()), "siss_wmse_bwd")
    _count()
    return grad


# ------------------------------------------------------------------------------------------------
# plain squared error
# ------------------------------------------------------------------------------------------------
def _promoted(pred: torch.Tensor, target: torch.Tensor) -> torch.dtype:
    return torch.promote_types(pred.dtype, target.dtype)


def sqerr_fwd(pred: torch.Tensor, target: torch.Tensor, alpha: Optional[float] = None):
    """loss = (pred - target)^2 [, scaled = alpha * loss] in the promoted dtype."""
    _need_cuda(pred, target)
    if pred.shape != target.shape:
        raise ValueError("pred and target must share one shape")
    pred, target = _c(pred), _c(target)
    od = _promoted(pred, target)
    loss = torch.empty(pred.shape, dtype=od, device=pred.device)
    scaled = torch.empty_like(loss) if alpha is not None else None
    _lib.check(_lib.load().siss_sqerr_fwd(_ptr(pred), _dt(pred), _ptr(target), _dt(target), _ptr(loss), _ptr(scaled),
                                          float(alpha or 0.0), pred.numel(), _stream()), "siss_sqerr_fwd")
    _count()
    return loss, scaled


def sqerr_bwd(pred, target, go_loss=None, go_scaled=None, alpha: float = 0.0):
    _need_cuda(pred, target)
    pred, target = _c(pred), _c(target)
    od = _promoted(pred, target)
    g1, s1 = _grad_out(go_loss, pred.shape, od)
    g2, s2 = _grad_out(go_scaled, pred.shape, od)
    grad = torch.empty_like(pred)
    if g1 is None and g2 is None:
        return grad.zero_()
    _lib.check(_lib.load().siss_sqerr_bwd(_ptr(pred), _dt(pred), _ptr(target), _dt(target), _ptr(g1), s1, _ptr(g2), s2,
                                          float(alpha), _DTYPES[od], _ptr(grad), pred.numel(), _stream()),
               "siss_sqerr_bwd")
    _count()
    return grad


def dual_mse_fwd_bwd(pred_x, pred_a, target_x, target_a, go_x: float, go_a: float):
    """No-IS / EraseDiff fast path. Returns (grad_x, grad_a, row_loss_x, row_loss_a)."""
    if BINDING == "torch":
        out = _ext_for(pred_x, pred_a, target_x, target_a).dual_mse_fwd_bwd(pred_x, pred_a, target_x, target_a, float(go_x),
                                                                            float(go_a))
        _count()
        return out
    dev = _need_cuda(pred_x, pred_a, target_x, target_a)
    if not (pred_x.shape == pred_a.shape == target_x.shape == target_a.shape):
        raise ValueError("preds and targets must share one shape")
    if pred_x.dtype != pred_a.dtype or target_x.dtype != target_a.dtype:
        raise ValueError("pred_x/pred_a and target_x/target_a must pairwise share dtypes")
    same = target_a is target_x or (target_a.data_ptr() == target_x.data_ptr())
    pred_x, pred_a, target_x = _c(pred_x), _c(pred_a), _c(target_x)
    target_a = target_x if same else _c(target_a)
    B, D = _rows(pred_x)
    grad_x, grad_a = torch.empty_like(pred_x), torch.empty_like(pred_a)
    rows = torch.empty((2, B), dtype=torch.float32, device=dev)
    ws = _row_workspace(dev, B)
    _lib.check(_lib.load().siss_dual_mse_fwd_bwd(
        _ptr(pred_x), _ptr(pred_a), _dt(pred_x), _ptr(target_x), _ptr(target_a), _dt(target_x), float(go_x),
        float(go_a), _ptr(grad_x), _ptr(grad_a), _ptr(rows[0]), _ptr(rows[1]), _ptr(ws), B, D, _stream()),
        "siss_dual_mse_fwd_bwd")
    _count()
    return grad_x, grad_a, rows[0], rows[1]


def dual_mse_rng_fwd_bwd(pred_x, pred_a, target_x, go_x: float, go_a: float, seed: int, draw: int, elem_offset: int = 0,
                         d_draw: Optional[torch.Tensor] = None, want_target: bool = False):
    """EraseDiff fast path with the uniform forget target drawn in-kernel from the counter-based stream (opt-in,
    siss_b200.rng). Returns (grad_x, grad_a, row_loss_x, row_loss_a, target_a-or-None)."""
    dev = _need_cuda(pred_x, pred_a, target_x)
    if not (pred_x.shape == pred_a.shape == target_x.shape) or pred_x.dtype != pred_a.dtype:
        raise ValueError("pred_x, pred_a and target_x must share one shape; preds share a dtype")
    pred_x, pred_a, target_x = _c(pred_x), _c(pred_a), _c(target_x)
    B, D = _rows(pred_x)
    grad_x, grad_a = torch.empty_like(pred_x), torch.empty_like(pred_a)
    tgt = torch.empty_like(pred_a) if want_target else None
    rows = torch.empty((2, B), dtype=torch.float32, device=dev)
    if pred_x.numel() == 0:
        return grad_x, grad_a, rows[0].zero_(), rows[1].zero_(), tgt
    ws = _row_workspace(dev, B)
    _lib.check(_lib.load().siss_dual_mse_rng_fwd_bwd(
        _ptr(pred_x), _ptr(pred_a), _dt(pred_x), _ptr(target_x), _dt(target_x), int(seed) & 0xFFFFFFFFFFFFFFFF,
        int(draw), _ptr(d_draw), int(elem_offset), float(go_x), float(go_a), _ptr(grad_x), _ptr(grad_a), _ptr(tgt),
        _ptr(rows[0]), _ptr(rows[1]), _ptr(ws), B, D, _stream()), "siss_dual_mse_rng_fwd_bwd")
    _count()
    return grad_x, grad_a, rows[0], rows[1], tgt


# ------------------------------------------------------------------------------------------------
# K4
# ------------------------------------------------------------------------------------------------
def norm3(g_x: torch.Tensor, g_a: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{sum g_x^2, sum g_a^2, sum g_x g_a} as a float64[3] device tensor (K4a)."""
    if BINDING == "torch":
        ext = _ext_for(g_x, g_a)
        if out is None:
            out = torch.empty(3, dtype=torch.float64, device=g_x.device)
        ext.norm3_(g_x, g_a, out)
        _count()
        return out
    dev = _need_cuda(g_x, g_a)
    if g_x.dtype != torch.float32 or g_a.dtype != torch.float32:
        raise SissLibraryError("gradient buffers must be float32")
    if g_x.numel() != g_a.numel() or not g_x.is_contiguous() or not g_a.is_contiguous():
        raise ValueError("g_x and g_a must be contiguous and equally sized")
    if out is None:
        out = torch.empty(3, dtype=torch.float64, device=dev)
    _lib.check(_lib.load().siss_norm3(_ptr(g_x), _ptr(g_a), g_x.numel(), _ptr(out), _ptr(_norm_workspace(dev)),
                                      _stream()), "siss_norm3")
    _count()
    return out


def combine(g_x: torch.Tensor, g_a: torch.Tensor, sums3: torch.Tensor, mode: int, value: float,
            max_norm: float = 1.0, inf_guard: bool = False, out: Optional[torch.Tensor] = None,
            stats: Optional[torch.Tensor] = None):
    """out = clip * (g_x - s g_a) (K4b). Returns (out, stats5) with stats5 = [||g_x||, ||g_a||, s,
    ||g_x - s g_a||, clip] on the device."""
    if BINDING == "torch":
        ext = _ext_for(g_x, g_a, sums3)
        if out is None:
            out = torch.empty_like(g_x)
        if stats is None:
            stats = torch.empty(5, dtype=torch.float32, device=g_x.device)
        ext.combine_(g_x, g_a, sums3, int(mode), float(value), float(max_norm if max_norm is not None else 0.0),
                     bool(inf_guard), out, stats)
        _count()
        return out, stats
    dev = _need_cuda(g_x, g_a, sums3)
    if g_x.dtype != torch.float32 or g_a.dtype != torch.float32 or sums3.dtype != torch.float64:
        raise SissLibraryError("gradient buffers must be float32 and sums3 float64")
    if out is None:
        out = torch.empty_like(g_x)
    if stats is None:
        stats = torch.empty(5, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().siss_combine(_ptr(g_x), _ptr(g_a), _ptr(out), g_x.numel(), _ptr(sums3), int(mode),
                                        float(value), float(max_norm if max_norm is not None else 0.0),
                                        int(bool(inf_guard)), _ptr(stats), _stream()), "siss_combine")
    _count()
    return out, stats


# ------------------------------------------------------------------------------------------------
# fused statistics epilogue
# ------------------------------------------------------------------------------------------------
STAT_KEYS = tuple(f"{q}/{s}" for q in ("loss_x", "loss_a", "importance_weight_x", "importance_weight_a")
                  for s in ("mean", "max", "min", "std"))


def batch_stats(row_loss_x: Optional[torch.Tensor], row_loss_a: Optional[torch.Tensor], w_x: Optional[torch.Tensor],
                w_a: Optional[torch.Tensor], elems_per_sample: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The 16 logging scalars of delete_celeb.py:626-656 in one launch (order: ``STAT_KEYS``).
    Inputs are the per-sample sums / weights K2 and K3 return; any may be None (-> NaN outputs)."""
    if BINDING == "torch":
        present = [t for t in (row_loss_x, row_loss_a, w_x, w_a) if t is not None]
        if not present:
            raise ValueError("batch_stats needs at least one input")
        ext = _ext_for(*present)
        if out is None:
            out = torch.empty(16, dtype=torch.float32, device=present[0].device)
        ext.batch_stats_(row_loss_x, row_loss_a, w_x, w_a, int(elems_per_sample), out)
        _count()
        return out
    present = [t for t in (row_loss_x, row_loss_a, w_x, w_a) if t is not None]
    if not present:
        raise ValueError("batch_stats needs at least one input")
    dev = _need_cuda(*present)
    B = present[0].numel()
    for t in present:
        if t.dtype != torch.float32 or t.numel() != B or not t.is_contiguous():
            raise ValueError("batch_stats inputs must be contiguous float32 vectors of one length")
    if out is None:
        out = torch.empty(16, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().siss_batch_stats(_ptr(row_loss_x), _ptr(row_loss_a), _ptr(w_x), _ptr(w_a), B,
                                            int(elems_per_sample), _ptr(out), _stream()), "siss_batch_stats")
    _count()
    return out


# ------------------------------------------------------------------------------------------------
# multi-tensor K4 (per-parameter gradient tensors, no flat buffers)
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# Membership-loss metric (metrics/class_membership.py:66-116)
# ------------------------------------------------------------------------------------------------
def membership_add_noise(x0: torch.Tensor, a0: torch.Tensor, noise: torch.Tensor, timestep: int,
                         alphas_cumprod: torch.Tensor, row0: int, rows: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Noisy all / deletion batches for expanded rows [row0, row0+rows) of the I x N_n grid (image r // N_n, noise
    r % N_n) without materialising the expansion. x0, a0: [I, ...]; noise: [N_n, ...]."""
    dev = _need_cuda(x0, a0, noise)
    if x0.shape != a0.shape or x0.shape[1:] != noise.shape[1:] or not (x0.dtype == a0.dtype == noise.dtype):
        raise ValueError("x0 / a0 [I, ...] and noise [N_n, ...] must share trailing shape and dtype")
    x0, a0, noise = _c(x0), _c(a0), _c(noise)
    I, D = _rows(x0)
    n_noise = noise.shape[0]
    if row0 < 0 or rows < 0 or row0 + rows > I * n_noise:
        raise ValueError(f"rows [{row0}, {row0 + rows}) outside the {I} x {n_noise} expansion")
    ac = _table(alphas_cumprod, dev)
    if not 0 <= int(timestep) < ac.numel():
        raise IndexError(f"timestep {timestep} outside the {ac.numel()}-step schedule")
    shape = (rows,) + tuple(x0.shape[1:])
    xt_x, xt_a = torch.empty(shape, dtype=x0.dtype, device=dev), torch.empty(shape, dtype=x0.dtype, device=dev)
    if xt_x.numel() == 0:
        return xt_x, xt_a
    _lib.check(_lib.load().siss_membership_add_noise(_ptr(x0), _ptr(a0), _ptr(noise), _ptr(ac), ac.numel(), int(timestep),
                                                     _ptr(xt_x), _ptr(xt_a), int(row0), int(rows), n_noise, D, _dt(x0),
                                                     _stream()), "siss_membership_add_noise")
    _count()
    return xt_x, xt_a


def membership_sqerr(pred_x: torch.Tensor, pred_a: torch.Tensor, noise: torch.Tensor,
                     row0: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-row sum over C,H,W of (pred - noise[(row0 + q) % N_n])^2 for both predictions, one pass. -> 2x [rows] fp32."""
    dev = _need_cuda(pred_x, pred_a, noise)
    if pred_x.shape != pred_a.shape or pred_x.shape[1:] != noise.shape[1:]:
        raise ValueError("pred_x / pred_a [rows, ...] and noise [N_n, ...] must share trailing shape")
    pred_x, pred_a, noise = _c(pred_x.float()), _c(pred_a.float()), _c(noise)
    rows, D = _rows(pred_x)
    sum_x = torch.empty(rows, dtype=torch.float32, device=dev)
    sum_a = torch.empty(rows, dtype=torch.float32, device=dev)
    if pred_x.numel() == 0:
        return sum_x.zero_(), sum_a.zero_()
    ws = _row_workspace(dev, rows)
    _lib.check(_lib.load().siss_membership_sqerr(_ptr(pred_x), _ptr(pred_a), _ptr(noise), _dt(noise), _ptr(sum_x),
                                                 _ptr(sum_a), _ptr(ws), int(row0), rows, noise.shape[0], D, _stream()),
               "siss_membership_sqerr")
    _count()
    return sum_x, sum_a


class MultiTensorPlan:
    """Device-side pointer/size/chunk tables for ``siss_mt_norm3`` / ``siss_mt_combine`` over lists of fp32
    tensors ``accum_x[i]``, ``accum_a[i]`` (and ``out[i]``, default: write back into ``accum_x[i]``). Build
    once per set of tensors; the tables hold raw pointers, so the tensors must stay alive and in place."""

    def __init__(self, accum_x, accum_a, out=None):
        accum_x, accum_a = list(accum_x), list(accum_a)
        out = accum_x if out is None else list(out)
        if not accum_x or not (len(accum_x) == len(accum_a) == len(out)):
            raise ValueError("need equally long, non-empty tensor lists")
        self.dev = _need_cuda(*accum_x, *accum_a, *out)
        for x, a, o in zip(accum_x, accum_a, out):
            if not (x.dtype == a.dtype == o.dtype == torch.float32):
                raise SissLibraryError("multi-tensor combine handles float32 gradients")
            if not (x.numel() == a.numel() == o.numel()) or not (x.is_contiguous() and a.is_contiguous() and o.is_contiguous()):
                raise ValueError("tensors of one index must be contiguous and equally sized")
        self.keep = (accum_x, accum_a, out)
        chunk = _lib.load().siss_mt_chunk_elems()
        sizes = [x.numel() for x in accum_x]
        prefix = [0]
        for n in sizes:
            prefix.append(prefix[-1] + (n + chunk - 1) // chunk)
        i64 = dict(dtype=torch.int64, device=self.dev)
        self.n = len(sizes)
        self.total_chunks = prefix[-1]
        self.gx = torch.tensor([t.data_ptr() for t in accum_x], **i64)
        self.ga = torch.tensor([t.data_ptr() for t in accum_a], **i64)
        self.out = torch.tensor([t.data_ptr() for t in out], **i64)
        self.sizes = torch.tensor(sizes, **i64)
        self.prefix = torch.tensor(prefix, **i64)
        self.sums3 = torch.zeros(3, dtype=torch.float64, device=self.dev)
        self.stats = torch.zeros(5, dtype=torch.float32, device=self.dev)

    def norm3(self) -> torch.Tensor:
        _lib.check(_lib.load().siss_mt_norm3(_ptr(self.gx), _ptr(self.ga), _ptr(self.sizes), _ptr(self.prefix), self.n,
                                             self.total_chunks, _ptr(self.sums3), _ptr(_norm_workspace(self.dev)),
                                             _stream()), "siss_mt_norm3")
        _count()
        return self.sums3

    def combine(self, mode: int, value: float, max_norm: float = 1.0, inf_guard: bool = False) -> torch.Tensor:
        """K4a + K4b over the list; results land in ``out[i]``. Returns the device stats5 tensor."""
        self.norm3()
        _lib.check(_lib.load().siss_mt_combine(_ptr(self.gx), _ptr(self.ga), _ptr(self.out), _ptr(self.sizes),
                                               _ptr(self.prefix), self.n, self.total_chunks, _ptr(self.sums3), int(mode),
                                               float(value), float(max_norm if max_norm is not None else 0.0),
                                               int(bool(inf_guard)), _ptr(self.stats), _stream()), "siss_mt_combine")
        _count()
        return self.stats
