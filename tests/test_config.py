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
