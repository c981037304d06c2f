"""tools/unet2d.py — the plain-PyTorch restatement of diffusers' UNet2DModel that bench.py's steps/s block runs the path
around (measurement infrastructure; diffusers is not installed). What pins the restatement: the parameter count of the
google/ddpm-celebahq-256 configuration is the published 113 673 219 — any wrong block, channel plan or bias changes it —
and every parameter takes part in a forward / backward pass with the UNet2DModel call convention of the task loops."""
import sys
from pathlib import Path

import pytest
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import unet2d  # noqa: E402


def test_celebahq256_parameter_count_is_the_checkpoints():
    with torch.device("meta"):
        m = unet2d.celebahq256()
    assert sum(p.numel() for p in m.parameters()) == unet2d.CELEBAHQ256_PARAMS == 113_673_219
    # 248.5 G multiply-adds per 256x256 sample
    assert abs(unet2d.forward_flops(m, 256, 3) / 1e9 - 497.03) < 0.05


def test_tshirt28_matches_the_reference_config():
    """config/train_tshirt_mnist.yaml:25-41: (64, 128, 256), attention in the middle pair; SURVEY's estimate was ~15 M."""
    with torch.device("meta"):
        m = unet2d.tshirt28()
    assert sum(p.numel() for p in m.parameters()) == 14_735_745


@pytest.mark.parametrize("make,ch,res", [(unet2d.tshirt28, 1, 28), (unet2d.celebahq256, 3, 32)])
def test_forward_backward_reaches_every_parameter(make, ch, res):
    torch.manual_seed(0)
    m = make()
    x = torch.randn(2, ch, res, res)
    t = torch.tensor([999, 3])
    y = m(x, t, return_dict=False)[0]
    assert y.shape == x.shape and y.dtype == torch.float32 and torch.isfinite(y).all()
    (y ** 2).mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters())
    # the time embedding matters: another t changes the output of the same input row
    y2 = m(x, torch.tensor([999, 500]), return_dict=False)[0]
    assert torch.equal(y[0], y2[0]) and not torch.equal(y[1], y2[1])
