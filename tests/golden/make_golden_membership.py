"""Generate tests/golden/membership_*.npz by executing the REFERENCE's own metric class
(/root/reference/metrics/class_membership.py::MembershipLoss). Build container only:

    python tests/golden/make_golden_membership.py

The reference class is run unmodified on CPU with seeded list datasets, a 2-parameter stub UNet and a scheduler
stub whose ``add_noise`` is the plain-torch restatement of diffusers 0.27.2's (diffusers is not installed; same
expression as make_golden.py). Stored: the sampled images, the noise, the timesteps and the losses it returned.
"""
import importlib.util
import random
from pathlib import Path

import numpy as np
import torch

from make_golden import add_noise, alphas_cumprod

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference/metrics/class_membership.py")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_class_membership", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.MembershipLoss


class SchedulerStub:
    def __init__(self, schedule):
        self.alphas_cumprod = alphas_cumprod(schedule)

    def add_noise(self, x0, noise, t):
        return add_noise(self.alphas_cumprod, x0, noise, t)


class StubUNet(torch.nn.Module):
    def forward(self, x, timesteps, return_dict=False, **kw):
        # depends on the timestep so a wrong / missing timestep tensor shows up in the result
        return (x * 0.75 + 0.05 + timesteps.reshape(-1, 1, 1, 1).float() * 1e-4,)


CASES = [
    dict(name="membership_tshirt", shape=(1, 28, 28), n_all=40, n_del=1, images=6, noises=5, eval_bs=8,
         timesteps=[250, 999], schedule="linear", seed=46),
    dict(name="membership_rgb_ragged", shape=(3, 10, 9), n_all=12, n_del=7, images=5, noises=3, eval_bs=4,
         timesteps=[0, 500], schedule="linear", seed=42),
]


def main():
    Ref = load_reference()
    for c in CASES:
        torch.manual_seed(c["seed"]); random.seed(c["seed"])
        ds_all = [torch.rand(c["shape"]) * 2 - 1 for _ in range(c["n_all"])]
        ds_del = [torch.rand(c["shape"]) * 2 - 1 for _ in range(c["n_del"])]
        m = Ref(ds_all, ds_del, SchedulerStub(c["schedule"]), StubUNet(), c["images"], c["noises"], c["eval_bs"], "cpu")
        m.sample_images()
        m.sample_noises()
        out = m.compute_membership_losses(c["timesteps"])
        data = dict(all_images=m.all_sampled_images.numpy(), deletion_images=m.deletion_sampled_images.numpy(),
                    noise=m.noise.numpy(), timesteps=np.array(c["timesteps"]), eval_bs=np.array(c["eval_bs"]),
                    alphas_cumprod=m.noise_scheduler.alphas_cumprod.numpy(),
                    losses=np.array([[float(a), float(d)] for a, d in out], dtype=np.float64),
                    losses_f32=np.array([[a.item(), d.item()] for a, d in out], dtype=np.float32),
                    dataset_all=torch.stack(ds_all).numpy(), dataset_deletion=torch.stack(ds_del).numpy(),
                    seed=np.array(c["seed"]))
        np.savez_compressed(HERE / f"{c['name']}.npz", **data)
        print(c["name"], data["losses"].tolist())


if __name__ == "__main__":
    main()
