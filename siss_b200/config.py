"""Reader for the reference's Hydra config keys (the ones the hot path consumes).

The reference is driven by ``python main.py --config-name=delete_{tshirt,celeb,sd}`` (Hydra + OmegaConf,
neither installed in this image). The hot path reads exactly these keys:

    deletion.loss_fn, deletion.loss_params.{lambd, superfactor}, deletion.superfactor_decay,
    deletion.scaling_norm, deletion.eta, train_batch_size, gradient_accumulation_steps, mixed_precision,
    optimizer.{lr, betas, weight_decay, eps}, random_seed, scheduler.* (beta schedule)

``load_config`` parses such a YAML file with PyYAML, honouring Hydra's ``defaults:`` list the way the
reference uses it (``delete_tshirt.yaml`` inherits ``train_tshirt_mnist.yaml``; ``_self_`` marks where the
file's own keys are merged), and ``step_from_config`` turns it into an :class:`siss_b200.step.UnlearnStep`
with the same semantics the task loops give those keys. Key names are the contract; nothing is renamed.
"""
from __future__ import annotations

import os
from typing import Any, Dict, Optional

import yaml

TWO_TERM = ("importance_sampling_with_mixture", "double_forward_with_neg_del", "erasediff")


def _merge(base: Dict[str, Any], over: Dict[str, Any]) -> Dict[str, Any]:
    out = dict(base)
    for k, v in over.items():
        if isinstance(v, dict) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def load_config(path: str) -> Dict[str, Any]:
    """YAML -> nested dict, resolving a top-level Hydra ``defaults:`` list relative to the file's directory."""
    with open(path) as f:
        own = yaml.safe_load(f) or {}
    defaults = own.pop("defaults", None)
    if not defaults:
        return own
    merged: Dict[str, Any] = {}
    self_done = False
    for entry in defaults:
        if entry == "_self_":
            merged = _merge(merged, own)
            self_done = True
        elif isinstance(entry, str):
            merged = _merge(merged, load_config(os.path.join(os.path.dirname(path), entry + ".yaml")))
        else:
            raise ValueError(f"unsupported defaults entry {entry!r} (only plain config names and _self_)")
    if not self_done:                      # Hydra >= 1.1: _self_ last when absent
        merged = _merge(merged, own)
    return merged


def hot_path_params(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """The hot-path subset of a loaded config, with the reference's defaults and validation."""
    d = cfg.get("deletion", {}) or {}
    loss_fn = d.get("loss_fn")
    if loss_fn is None:
        raise KeyError("deletion.loss_fn is required")
    lp = d.get("loss_params") or {}
    out = dict(
        loss_fn=loss_fn,
        lambd=lp.get("lambd"),
        superfactor=lp.get("superfactor"),
        superfactor_decay=d.get("superfactor_decay"),
        scaling_norm=d.get("scaling_norm"),
        eta=d.get("eta"),
        train_batch_size=int(cfg.get("train_batch_size", 1)),
        gradient_accumulation_steps=int(cfg.get("gradient_accumulation_steps", 1)),
        mixed_precision=cfg.get("mixed_precision"),
        random_seed=cfg.get("random_seed"),
    )
    for k in ("scaling_norm", "eta", "lambd", "superfactor"):
        if out[k] is not None:
            out[k] = float(out[k])          # YAML 1.1 reads `1e-3` as a string
    return out


def step_from_config(cfg: Dict[str, Any], unet, scheduler, combiner, max_norm: float = 1.0,
                     inf_guard: bool = False, superfactor_decay_on: str = "micro_step"):
    """UnlearnStep with the config's meaning: ``scaling_norm`` applies to SISS / No-IS, ``eta`` to EraseDiff
    (delete_celeb.py:740-746), ``lambd`` / ``superfactor`` are the method kwargs (``**loss_params``, :622),
    the loss is divided by ``train_batch_size`` and by ``gradient_accumulation_steps`` (:686-691).
    ``superfactor_decay_on``: "micro_step" is what delete_celeb.py / delete_tshirt.py do with ``deletion.superfactor_decay``,
    "sync_step" what delete_sd.py does (once per optimiser step, delete_sd.py:1173-1193)."""
    from .step import UnlearnStep
    hp = hot_path_params(cfg)
    fn = hp["loss_fn"]
    return UnlearnStep(unet, scheduler, combiner, loss_fn=fn, train_batch_size=hp["train_batch_size"],
                       gradient_accumulation_steps=hp["gradient_accumulation_steps"], lambd=hp["lambd"],
                       superfactor=hp["superfactor"],
                       scaling_norm=hp["scaling_norm"] if fn in ("importance_sampling_with_mixture",
                                                                 "double_forward_with_neg_del") else None,
                       eta=hp["eta"] if fn == "erasediff" else None, max_norm=max_norm, inf_guard=inf_guard,
                       superfactor_decay=hp["superfactor_decay"], superfactor_decay_on=superfactor_decay_on)


def adamw_kwargs(cfg: Dict[str, Any]) -> Dict[str, Any]:
    """torch.optim.AdamW keyword arguments from ``optimizer:`` (config/delete_celeb.yaml:127-134)."""
    o = cfg.get("optimizer", {}) or {}
    kw: Dict[str, Any] = {}
    if "lr" in o:
        kw["lr"] = float(o["lr"])
    if "betas" in o:
        kw["betas"] = (float(o["betas"][0]), float(o["betas"][1]))
    if "weight_decay" in o:
        kw["weight_decay"] = float(o["weight_decay"])
    if "eps" in o:
        kw["eps"] = float(o["eps"])
    return kw


def ema_kwargs(cfg: Dict[str, Any]) -> Optional[Dict[str, Any]]:
    """``FusedCombineAdamW(ema=...)`` argument from the ``ema:`` block (config/train_tshirt_mnist.yaml:93-97;
    ``use_ema: false`` in the three deletion configs, e.g. config/delete_celeb.yaml:86): None when EMA is off, else
    the warm-up schedule the reference's trainer builds its EMAModel with (keys ema_max_decay / ema_inv_gamma /
    ema_power)."""
    e = cfg.get("ema")
    if not isinstance(e, dict):
        e = {"use_ema": cfg.get("use_ema", False)}          # delete_sd.yaml:75 keeps the flag at top level
    if str(e.get("use_ema", False)).strip().lower() not in ("true", "1", "yes"):
        return None
    kw: Dict[str, Any] = {"use_ema_warmup": True}
    if "ema_max_decay" in e:
        kw["decay"] = float(e["ema_max_decay"])
    if "ema_inv_gamma" in e:
        kw["inv_gamma"] = float(e["ema_inv_gamma"])
    if "ema_power" in e:
        kw["power"] = float(e["ema_power"])
    return kw
