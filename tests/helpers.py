"""Shared test helpers: golden-fixture loading and tolerance bookkeeping."""
from __future__ import annotations

from pathlib import Path
from typing import Dict

import numpy as np
import torch

GOLDEN_DIR = Path(__file__).resolve().parent / "golden"
GOLDEN_CASES = sorted(p.stem for p in GOLDEN_DIR.glob("*.npz"))

_LATENT_KEYS = ("x0", "a0", "noise", "xt_x", "xt_a")


def load_golden(name: str, device="cpu") -> Dict[str, object]:
    """Fixture -> dict of torch tensors; latents are restored to the dtype the reference saw."""
    z = np.load(GOLDEN_DIR / f"{name}.npz", allow_pickle=False)
    dt = getattr(torch, str(z["dtype"]))
    out: Dict[str, object] = {"dtype": dt, "lambd": float(z["lambd"]), "seed": int(z["seed"]),
                              "schedule": str(z["schedule"])}
    for k in z.files:
        if k in out or k == "dtype":
            continue
        v = z[k]
        if v.dtype.kind in "US":
            out[k] = str(v)
            continue
        if v.shape == () and v.dtype.kind in "iu":
            out[k] = int(v)
            continue
        t = torch.from_numpy(np.array(v))
        if k in _LATENT_KEYS:
            t = t.to(dt)
        out[k] = t.to(device)
    return out


def conditioning_of(case: Dict[str, object]) -> Dict[str, torch.Tensor]:
    if "encoder_hidden_states" in case:
        return {"encoder_hidden_states": case["encoder_hidden_states"]}
    return {}


def weight_tolerance(d_x: torch.Tensor, d_a: torch.Tensor, k: float = 8.0) -> torch.Tensor:
    """Relative tolerance on an importance weight whose exponent difference carries fp32 summation
    noise: |delta(d_x - d_a)| <= k * eps32 * (d_x + d_a); d(log w) <= |delta|. Plus 1e-5 floor."""
    eps = torch.finfo(torch.float32).eps
    return k * eps * (d_x.abs() + d_a.abs()).double() + 1e-5
