"""Membership-loss metric on the CUDA kernels — mirror of the reference's ``metrics/class_membership.py``
(SURVEY.md §8f rank 3): same class name, constructor arguments, ``sample_images`` / ``sample_noises`` /
``compute_membership_losses(timesteps)`` methods and return structure (a list of ``[all_loss, deletion_loss]``
0-dim tensors per timestep).

What changes is the data movement. The reference expands images and noise to ``[I*N_n, C, H, W]`` (:76-86), runs
``add_noise`` over both expansions (:92-93) and per eval batch materialises ``(pred - noise)**2`` before reducing
(:108-109). Here the expansion is an index map inside two kernels (``siss_membership_add_noise``,
``siss_membership_sqerr``): per eval batch, 3 reads + 2 writes per element to build both noisy batches and 3 reads
per element for both row sums; nothing of expanded size is ever allocated.
"""
from __future__ import annotations

import random
from typing import List

import torch

from .. import ops


class MembershipLoss:
    def __init__(self, dataset_all, dataset_deletion, noise_scheduler, unet, num_image_samples, num_noise_samples,
                 eval_batch_size, device):
        self.dataset_all = dataset_all
        self.dataset_deletion = dataset_deletion
        self.noise_scheduler = noise_scheduler          # anything with .alphas_cumprod (SissDDPMScheduler, diffusers)
        self.unet = unet
        self.num_image_samples = num_image_samples
        self.num_noise_samples = num_noise_samples
        self.eval_batch_size = eval_batch_size
        self.device = device

    def sample_images(self):
        """Same draws, in the same order, from Python's ``random`` as class_membership.py:30-62."""
        n_all, n_del = len(self.dataset_all), len(self.dataset_deletion)
        all_idx = random.sample(range(n_all), self.num_image_samples)
        if n_del == 1:
            del_idx = [0] * self.num_image_samples
        else:
            del_idx = random.sample(range(n_del), self.num_image_samples)
        self.all_sampled_images = torch.stack([self.dataset_all[i] for i in all_idx], dim=0).to(self.device)
        self.deletion_sampled_images = torch.stack([self.dataset_deletion[i] for i in del_idx], dim=0).to(self.device)

    def sample_noises(self):
        """class_membership.py:64-67 (sample images first)."""
        self.noise = torch.randn((self.num_noise_samples, *self.all_sampled_images.shape[1:]), device=self.device)

    @torch.no_grad()
    def compute_membership_losses(self, timesteps: List[int]):
        """class_membership.py:69-128: per timestep, mean over the I x N_n grid of the summed squared error of the
        UNet's noise prediction, for the sampled 'all' and 'deletion' images."""
        assert self.all_sampled_images.shape == self.deletion_sampled_images.shape
        total = self.all_sampled_images.shape[0] * self.num_noise_samples
        ac = self.noise_scheduler.alphas_cumprod
        losses = []
        for timestep in timesteps:
            all_rows, del_rows = [], []
            for r0 in range(0, total, self.eval_batch_size):
                rows = min(self.eval_batch_size, total - r0)
                xt_all, xt_del = ops.membership_add_noise(self.all_sampled_images, self.deletion_sampled_images,
                                                          self.noise, timestep, ac, r0, rows)
                ts = torch.full((rows,), timestep, device=self.device)
                out_all = self.unet(xt_all, ts, return_dict=False)[0]
                out_del = self.unet(xt_del, ts, return_dict=False)[0]
                s_all, s_del = ops.membership_sqerr(out_all, out_del, self.noise, r0)
                all_rows.append(s_all)
                del_rows.append(s_del)
            losses.append([torch.mean(torch.cat(all_rows)), torch.mean(torch.cat(del_rows))])
        return losses
