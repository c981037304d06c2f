"""Architecture-equivalent stand-in for ``diffusers.UNet2DModel`` in plain PyTorch — MEASUREMENT INFRASTRUCTURE ONLY.

diffusers is not installed in this image and the UNet is outside the hot path ("stays on diffusers"),
but the second half of BASELINE's metric — unlearning optimiser steps/s at 1 / 2 / 4 / 8 GPUs — needs a UNet with the real
FLOPs, activations and parameter layout around the path. This module restates the published architecture of
diffusers 0.27.2's UNet2DModel (environment.yml:232; source not on disk) for the two configurations the reference uses:

  * ``celebahq256()``  google/ddpm-celebahq-256, loaded by delete_celeb.py:181-186 (config/delete_celeb.yaml:124-125):
    block_out_channels (128, 128, 256, 256, 512, 512), 2 layers per block, self-attention in the 16x16 down / up block
    and the mid block, GroupNorm(32, eps 1e-6), SiLU, sinusoidal time embedding 128 -> 512.
    The restatement has exactly 113 673 219 parameters — the published count of that checkpoint — which is the check
    that the block structure is right (tests/test_unet2d_arch.py).
  * ``tshirt28()``     config/train_tshirt_mnist.yaml:25-41: (64, 128, 256), attention in the middle block pair.

Weights are random: the timing does not depend on them. Nothing in ``siss_b200/`` imports this file."""
from __future__ import annotations

import math
from typing import Sequence

import torch
from torch import nn
from torch.nn import functional as F


class ResBlock(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int, groups: int, eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x, emb):
        h = self.conv1(F.silu(self.norm1(x)))
        h = h + self.time_emb_proj(F.silu(emb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class SelfAttention(nn.Module):
    """One head over the H*W positions (attention_head_dim = channels, as the celebahq config resolves it), residual."""

    def __init__(self, ch: int, groups: int, eps: float, head_dim: int | None = None):
        super().__init__()
        self.heads = 1 if head_dim is None else max(ch // head_dim, 1)
        self.group_norm = nn.GroupNorm(groups, ch, eps=eps)
        self.to_q, self.to_k, self.to_v = nn.Linear(ch, ch), nn.Linear(ch, ch), nn.Linear(ch, ch)
        self.to_out = nn.Linear(ch, ch)

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x.reshape(b, c, h * w)).transpose(1, 2)                      # [B, HW, C]
        split = lambda t: t.reshape(b, h * w, self.heads, c // self.heads).transpose(1, 2)
        o = F.scaled_dot_product_attention(split(self.to_q(y)), split(self.to_k(y)), split(self.to_v(y)))
        o = self.to_out(o.transpose(1, 2).reshape(b, h * w, c))
        return x + o.transpose(1, 2).reshape(b, c, h, w)


class Down(nn.Module):
    def __init__(self, cin, cout, temb, layers, attn, downsample, groups, eps, pad):
        super().__init__()
        self.resnets = nn.ModuleList(ResBlock(cin if i == 0 else cout, cout, temb, groups, eps) for i in range(layers))
        self.attentions = nn.ModuleList(SelfAttention(cout, groups, eps) for _ in range(layers)) if attn else None
        self.pad = pad
        self.downsample = nn.Conv2d(cout, cout, 3, stride=2, padding=pad) if downsample else None

    def forward(self, x, emb):
        skips = []
        for i, r in enumerate(self.resnets):
            x = r(x, emb)
            if self.attentions is not None:
                x = self.attentions[i](x)
            skips.append(x)
        if self.downsample is not None:
            if self.pad == 0:
                x = F.pad(x, (0, 1, 0, 1))
            x = self.downsample(x)
            skips.append(x)
        return x, skips


class Up(nn.Module):
    def __init__(self, cin, cout, cprev, temb, layers, attn, upsample, groups, eps):
        super().__init__()
        self.resnets = nn.ModuleList(
            ResBlock((cprev if i == 0 else cout) + (cin if i == layers - 1 else cout), cout, temb, groups, eps)
            for i in range(layers))
        self.attentions = nn.ModuleList(SelfAttention(cout, groups, eps) for _ in range(layers)) if attn else None
        self.upsample = nn.Conv2d(cout, cout, 3, padding=1) if upsample else None

    def forward(self, x, skips, emb):
        for i, r in enumerate(self.resnets):
            x = r(torch.cat([x, skips.pop()], dim=1), emb)
            if self.attentions is not None:
                x = self.attentions[i](x)
        if self.upsample is not None:
            x = self.upsample(F.interpolate(x, scale_factor=2.0, mode="nearest"))
        return x


class UNet2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, block_out_channels: Sequence[int], down_attn: Sequence[bool],
                 up_attn: Sequence[bool], layers_per_block: int = 2, groups: int = 32, eps: float = 1e-5,
                 downsample_padding: int = 1, flip_sin_to_cos: bool = True, freq_shift: float = 0.0,
                 autocast_dtype: torch.dtype | None = torch.bfloat16):
        super().__init__()
        ch = list(block_out_channels)
        temb = 4 * ch[0]
        self.c0, self.flip, self.shift, self.autocast_dtype = ch[0], flip_sin_to_cos, freq_shift, autocast_dtype
        self.conv_in = nn.Conv2d(in_channels, ch[0], 3, padding=1)
        self.time_embedding = nn.Sequential(nn.Linear(ch[0], temb), nn.SiLU(), nn.Linear(temb, temb))
        self.down_blocks = nn.ModuleList()
        out = ch[0]
        for i, c in enumerate(ch):
            cin, out = out, c
            self.down_blocks.append(Down(cin, out, temb, layers_per_block, down_attn[i], i != len(ch) - 1, groups, eps,
                                         downsample_padding))
        self.mid_res1 = ResBlock(ch[-1], ch[-1], temb, groups, eps)
        self.mid_attn = SelfAttention(ch[-1], groups, eps)
        self.mid_res2 = ResBlock(ch[-1], ch[-1], temb, groups, eps)
        rev = ch[::-1]
        self.up_blocks = nn.ModuleList()
        out = rev[0]
        for i, c in enumerate(rev):
            prev, out = out, c
            cin = rev[min(i + 1, len(ch) - 1)]
            self.up_blocks.append(Up(cin, out, prev, temb, layers_per_block + 1, up_attn[i], i != len(ch) - 1, groups, eps))
        self.conv_norm_out = nn.GroupNorm(groups, ch[0], eps=eps)
        self.conv_out = nn.Conv2d(ch[0], out_channels, 3, padding=1)

    def _time_proj(self, t):
        half = self.c0 // 2
        freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / (half - self.shift))
        a = t.float()[:, None] * freqs[None]
        return torch.cat([a.cos(), a.sin()] if self.flip else [a.sin(), a.cos()], dim=-1)

    def _forward(self, x, timesteps):
        emb = self.time_embedding(self._time_proj(timesteps))
        x = self.conv_in(x)
        skips = [x]
        for blk in self.down_blocks:
            x, s = blk(x, emb)
            skips += s
        x = self.mid_res2(self.mid_attn(self.mid_res1(x, emb)), emb)
        for blk in self.up_blocks:
            x = blk(x, skips, emb)
        return self.conv_out(F.silu(self.conv_norm_out(x)))

    def forward(self, x, timesteps, return_dict=False, **kw):
        """UNet2DModel call convention of the task loops: ``unet(x_t, t, return_dict=False)[0]`` (losses/ddpm_deletion_loss.py:24);
        bf16 autocast with an fp32 result, as accelerate's mixed_precision wrapper returns it (delete_celeb.py:102-108)."""
        if timesteps.dim() == 0:
            timesteps = timesteps[None].expand(x.shape[0])
        if self.autocast_dtype is not None and x.is_cuda:
            with torch.autocast("cuda", dtype=self.autocast_dtype):
                y = self._forward(x, timesteps)
        else:
            y = self._forward(x.float(), timesteps)
        return (y.float(),)


CELEBAHQ256_PARAMS = 113_673_219


def celebahq256(**kw) -> UNet2D:
    """google/ddpm-celebahq-256 (UNet2DModel config of the checkpoint delete_celeb.py:181-186 loads)."""
    return UNet2D(3, 3, (128, 128, 256, 256, 512, 512), down_attn=(False, False, False, False, True, False),
                  up_attn=(False, True, False, False, False, False), layers_per_block=2, groups=32, eps=1e-6,
                  downsample_padding=0, flip_sin_to_cos=False, freq_shift=1.0, **kw)


def tshirt28(**kw) -> UNet2D:
    """config/train_tshirt_mnist.yaml:25-41 (UNet2DModel defaults for everything the yaml does not set)."""
    return UNet2D(1, 1, (64, 128, 256), down_attn=(False, True, False), up_attn=(False, True, False), layers_per_block=2,
                  groups=32, eps=1e-5, downsample_padding=1, flip_sin_to_cos=True, freq_shift=0.0, **kw)


def forward_flops(model: UNet2D, res: int, in_channels: int) -> float:
    """Multiply-add FLOPs (2 per MAC) of one forward pass for ONE sample: convolutions, linears and the two attention GEMMs,
    counted with forward hooks on a meta-device pass."""
    total = [0.0]

    def conv_hook(m, inp, out):
        total[0] += 2.0 * out.numel() * (m.in_channels // m.groups) * m.kernel_size[0] * m.kernel_size[1]

    def lin_hook(m, inp, out):
        total[0] += 2.0 * out.numel() * m.in_features

    def attn_hook(m, inp, out):
        b, c, h, w = inp[0].shape
        total[0] += 2.0 * 2.0 * b * (h * w) ** 2 * c

    hooks = []
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            hooks.append(m.register_forward_hook(conv_hook))
        elif isinstance(m, nn.Linear):
            hooks.append(m.register_forward_hook(lin_hook))
        elif isinstance(m, SelfAttention):
            hooks.append(m.register_forward_hook(attn_hook))
    dev = next(model.parameters()).device
    with torch.no_grad():
        model._forward(torch.zeros(1, in_channels, res, res, device=dev), torch.zeros(1, dtype=torch.long, device=dev))
    for h in hooks:
        h.remove()
    return total[0]
