"""Noise schedule + ``add_noise`` with the diffusers ``DDPMScheduler`` call surface.

The reference builds ``DDPMScheduler(...)`` from a saved pipeline and only touches four things on
the hot path: ``alphas_cumprod`` (delete_celeb.py:367-371), ``config.num_train_timesteps``
(:594) and ``add_noise(original_samples, noise, timesteps)`` (:602-603). diffusers==0.27.2 is
pinned by the reference's environment.yml:232 but is not installed here, so this class restates
that surface (same names, same argument order) on top of the fused K1 kernel.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Tuple

import torch

from . import ops


def make_betas(num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
               beta_schedule: str = "linear") -> torch.Tensor:
    """fp32 beta table, as diffusers 0.27.2 DDPMScheduler.__init__ builds it."""
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
    if beta_schedule == "scaled_linear":  # Stable Diffusion
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
    raise NotImplementedError(f"beta_schedule {beta_schedule!r}")


class SissDDPMScheduler:
    """Minimal DDPMScheduler stand-in whose ``add_noise`` runs the sm_100a kernel."""

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02,
                 beta_schedule: str = "linear", prediction_type: str = "epsilon"):
        self.betas = make_betas(num_train_timesteps, beta_start, beta_end, beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)  # fp32, CPU (as in diffusers)
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, beta_start=beta_start,
                                      beta_end=beta_end, beta_schedule=beta_schedule,
                                      prediction_type=prediction_type)
        self._dev_tables = {}

    def _ac(self, device: torch.device) -> torch.Tensor:
        t = self._dev_tables.get(device)
        if t is None:
            t = self.alphas_cumprod.to(device=device, dtype=torch.float32).contiguous()
            self._dev_tables[device] = t
        return t

    def gamma_sigma(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        """gamma = sqrt(abar), sigma = sqrt(1 - abar) as the tasks build them (delete_celeb.py:367-371)."""
        gamma = (self.alphas_cumprod ** 0.5).to(device)
        sigma = ((1 - self.alphas_cumprod) ** 0.5).to(device)
        return gamma, sigma

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor,
                  timesteps: torch.Tensor) -> torch.Tensor:
        return ops.add_noise(original_samples, noise, timesteps, self._ac(original_samples.device))

    def add_noise_pair(self, keep_samples: torch.Tensor, forget_samples: torch.Tensor, noise: torch.Tensor,
                       timesteps: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """Both ``add_noise`` calls of delete_celeb.py:602-603 in one launch."""
        return ops.add_noise_pair(keep_samples, forget_samples, noise, timesteps, self._ac(keep_samples.device))
