"""CPU: the Hydra-key reader (defaults inheritance, key names, value coercion) and, when the reference
checkout is present (build container only), that its three deletion configs parse to the values the
survey recorded."""
from pathlib import Path

import pytest

from siss_b200 import config as C

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference/config")


def test_example_config_inheritance_and_keys():
    cfg = C.load_config(str(ROOT / "examples" / "config" / "delete_tshirt_like.yaml"))
    hp = C.hot_path_params(cfg)
    assert hp["loss_fn"] == "importance_sampling_with_mixture" and hp["lambd"] == 0.5
    assert hp["scaling_norm"] == 5.0 and hp["eta"] == 1e-3            # `1e-3` is a YAML-1.1 string: coerced
    assert hp["train_batch_size"] == 64                                 # own file overrides the inherited 128
    assert hp["gradient_accumulation_steps"] == 1 and hp["mixed_precision"] is None and hp["random_seed"] == 46
    kw = C.adamw_kwargs(cfg)
    assert kw == {"lr": 5e-5, "betas": (0.95, 0.999), "weight_decay": 1e-6, "eps": 1e-8}   # lr overridden, rest inherited
    assert cfg["scheduler"]["beta_schedule"] == "linear"


def test_missing_loss_fn_is_an_error():
    with pytest.raises(KeyError):
        C.hot_path_params({"deletion": {}})


@pytest.mark.skipif(not REF.exists(), reason="reference checkout not present (GPU box)")
def test_reference_configs_parse():
    expect = {"delete_tshirt": dict(scaling_norm=5.0, eta=1e-3, lambd=0.5, train_batch_size=64, gradient_accumulation_steps=1),
              "delete_celeb": dict(scaling_norm=500.0, eta=None, lambd=0.5, train_batch_size=4, gradient_accumulation_steps=16),
              "delete_sd": dict(scaling_norm=750.0, eta=1e-2, lambd=0.5, train_batch_size=1, gradient_accumulation_steps=16)}
    for name, want in expect.items():
        hp = C.hot_path_params(C.load_config(str(REF / f"{name}.yaml")))
        assert hp["loss_fn"] == "importance_sampling_with_mixture"
        for k, v in want.items():
            assert hp[k] == v, (name, k, hp[k])


def test_ema_decay_schedule_matches_oracle_restatement():
    """siss_b200.optim.ema_decay_at vs the oracle's EMAModel.get_decay restatement, plus hand values."""
    import torch
    from oracle import siss_oracle as O
    from siss_b200.optim import ema_decay_at
    for kw in (dict(), dict(use_ema_warmup=True, inv_gamma=1.0, power=0.75, decay=0.9999),
               dict(decay=0.5, min_decay=0.1, update_after_step=3)):
        ema = O.OracleEMA([torch.zeros(1)], **kw)
        for step in range(0, 40):
            assert ema_decay_at(step, **kw) == ema.get_decay(step)
    assert ema_decay_at(1) == 0.0 and ema_decay_at(2) == 2.0 / 11.0
    assert abs(ema_decay_at(2, use_ema_warmup=True, inv_gamma=1.0, power=0.75) - (1 - 2 ** -0.75)) < 1e-15
    assert ema_decay_at(10 ** 9, decay=0.9999) == 0.9999


def test_ema_kwargs_from_config_keys():
    assert C.ema_kwargs({"ema": {"use_ema": False, "ema_max_decay": 0.9999}}) is None
    assert C.ema_kwargs({"use_ema": "False"}) is None and C.ema_kwargs({}) is None
    kw = C.ema_kwargs({"ema": {"use_ema": True, "ema_inv_gamma": 1.0, "ema_power": 0.75, "ema_max_decay": 0.9999}})
    assert kw == {"use_ema_warmup": True, "decay": 0.9999, "inv_gamma": 1.0, "power": 0.75}
    if REF.is_dir():      # the reference's trainer enables it, its three deletion configs turn it off
        assert C.ema_kwargs(C.load_config(str(REF / "train_tshirt_mnist.yaml")))["power"] == 0.75
        for name in ("delete_tshirt", "delete_celeb", "delete_sd"):
            assert C.ema_kwargs(C.load_config(str(REF / f"{name}.yaml"))) is None
