"""Generate tests/golden/step_*.npz — one full optimiser step of the REFERENCE, executed from its own code. Build
container only:

    python tests/golden/make_golden_step.py

Per case: G micro-steps of  [reference DDPMDeletionLoss method (imported from /root/reference/losses/ddpm_deletion_loss.py)
-> the reference's inline backward / gradient-bookkeeping / combine / clip block, read from the task file's source text and
executed (delete_celeb.py:682-767 or its twin in delete_tshirt.py, which adds the inf guard)]  on a small conv net.
Stored: the net's initial parameters, every micro-step's inputs including the replayed Bernoulli mask / EraseDiff uniform
target, the gradients the block leaves in param.grad after clipping, and the three scalars it logs to wandb.
Third-party pieces that are not installed are stubbed as in the live test: add_noise (plain-torch restatement of diffusers'),
Accelerator.backward (loss / G, .backward) and clip_grad_norm_ (torch's)."""
import sys
import textwrap
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent))
from helpers import GoldenStepNet  # noqa: E402
from make_golden import add_noise, alphas_cumprod, load_reference  # noqa: E402

TASKS = {"delete_celeb": Path("/root/reference/delete_celeb.py"), "delete_tshirt": Path("/root/reference/delete_tshirt.py")}


def combine_block(path):
    lines = path.read_text().splitlines()
    start = [i for i, l in enumerate(lines) if l.strip() == "if loss is not None:"][1]
    end = next(i for i in range(start, len(lines)) if "accelerator.clip_grad_norm_(unet.parameters(), 1.0)" in lines[i])
    body = textwrap.dedent("\n".join(lines[start:end + 1]))
    src = ("def block(self, accelerator, unet, wandb, global_step, loss, weighted_loss_x, weighted_loss_a, accum_loss_x, "
           "accum_loss_a, torch, print):\n" + textwrap.indent(body, "    ") + "\n    return accum_loss_x, accum_loss_a\n")
    ns = {}
    exec(compile(src, f"<combine block of {path.name}>", "exec"), ns)
    return ns["block"]


class Accelerator:
    def __init__(self, G):
        self.G, self.sync_gradients = G, False

    def backward(self, loss, **kw):
        (loss / self.G).backward(**kw)

    def clip_grad_norm_(self, params, max_norm):
        return torch.nn.utils.clip_grad_norm_(params, max_norm)


CASES = [
    dict(name="step_siss_G1_celeb", task="delete_celeb", loss_fn="importance_sampling_with_mixture", G=1, lambd=0.5, scaling_norm=5.0),
    dict(name="step_siss_G3_tshirt", task="delete_tshirt", loss_fn="importance_sampling_with_mixture", G=3, lambd=0.3, scaling_norm=5.0),
    dict(name="step_nois_G2_celeb", task="delete_celeb", loss_fn="double_forward_with_neg_del", G=2, scaling_norm=500.0),
    dict(name="step_erasediff_G2_celeb", task="delete_celeb", loss_fn="erasediff", G=2, eta=0.05),
    dict(name="step_naive_G2_celeb", task="delete_celeb", loss_fn="naive_del", G=2),
    dict(name="step_neggrad_G1_tshirt", task="delete_tshirt", loss_fn="simple_neg_del", G=1, superfactor=0.7),
    dict(name="step_subscore_G2_tshirt", task="delete_tshirt", loss_fn="subscore_bernoulli", G=2, lambd=0.4, scaling_norm=5.0),
]


def main():
    Ref = load_reference()
    ac = alphas_cumprod("linear")
    gamma, sigma = ac ** 0.5, (1 - ac) ** 0.5
    for ci, c in enumerate(CASES):
        B, G, fn = 4, c["G"], c["loss_fn"]
        block = combine_block(TASKS[c["task"]])
        net = GoldenStepNet()
        data = {f"param/{n}": p.detach().clone().numpy() for n, p in net.named_parameters()}
        cfg = SimpleNamespace(train_batch_size=B, deletion=SimpleNamespace(loss_fn=fn, scaling_norm=c.get("scaling_norm"),
                                                                            eta=c.get("eta")))
        loss_obj = Ref(gamma=gamma, sigma=sigma)
        acc, logged = Accelerator(G), []
        wandb = SimpleNamespace(log=lambda d, step=None: logged.append(d))
        accum_x, accum_a = {}, {}
        kwargs = {k: c[k] for k in ("lambd", "superfactor") if k in c}
        torch.manual_seed(4242 + ci)
        for k in range(G):
            x0, a0 = torch.rand(B, 1, 8, 8) * 2 - 1, torch.rand(B, 1, 8, 8) * 2 - 1
            noise, t = torch.randn(B, 1, 8, 8), torch.randint(300, 1000, (B,))
            all_d = {"og_latents": x0, "noisy_latents": add_noise(ac, x0, noise, t)}
            del_d = {"og_latents": a0, "noisy_latents": add_noise(ac, a0, noise, t)}
            draw_seed = 500 + k
            torch.manual_seed(draw_seed)
            items = getattr(loss_obj, fn)(net, t, noise, {}, all_d, del_d, **kwargs)
            torch.manual_seed(draw_seed)                       # replay the method's own RNG draw
            if fn in ("importance_sampling_with_mixture", "subscore_bernoulli"):
                data[f"keep/{k}"] = (torch.rand(B) > c["lambd"]).numpy()
            elif fn == "erasediff":
                data[f"forget_target/{k}"] = torch.rand_like(noise).numpy()
            for nm, v in (("x0", x0), ("a0", a0), ("noise", noise), ("t", t)):
                data[f"{nm}/{k}"] = v.numpy()
            acc.sync_gradients = k == G - 1
            torch.manual_seed(9000 + k)                        # decouple the next inputs from the method's draws
            accum_x, accum_a = block(SimpleNamespace(cfg=cfg), acc, net, wandb, 0, items[0], items[5], items[6], accum_x,
                                     accum_a, torch, lambda *a, **k: None)
        for n, p in net.named_parameters():
            data[f"grad/{n}"] = p.grad.detach().clone().numpy()
        if logged:
            data["norm_loss_x"] = np.float64(float(logged[0]["gradient/norm_loss_x"]))
            data["norm_loss_a"] = np.float64(float(logged[0]["gradient/norm_loss_a"]))
            data["scaling_factor"] = np.float64(float(logged[0]["gradient/scaling_factor"]))
        data["meta"] = np.array([fn, c["task"], str(G), str(c.get("lambd")), str(c.get("scaling_norm")), str(c.get("eta")),
                                 str(c.get("superfactor"))])
        np.savez_compressed(HERE / f"{c['name']}.npz", **data)
        gn = float(torch.sqrt(sum((p.grad ** 2).sum() for p in net.parameters())))
        print(c["name"], "final grad norm", round(gn, 6), {k: float(v) for k, v in (logged[0].items() if logged else [])})


if __name__ == "__main__":
    main()
