#!/usr/bin/env python
"""A/B timing of the two row kernels (K1oK2 siss_add_noise_mixture, K3 siss_wmse_fwd_bwd) at the headline shape
(B = 64 x 3x256x256, bf16 latents, fp32 eps_hat) under whatever SISS_* environment knobs are set.

Per-launch CUDA-event brackets exactly like bench.py's second timed region (one bracket per launch on a busy
stream, median of N after dropping the first two), rotating over enough buffer sets that every launch reads its
inputs from DRAM. Also times a plain D2D copy of the same bytes under the same bracket (the size floor).
Prints one JSON line; `--tag` names the variant."""
import argparse
import json
import os
import statistics
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from siss_b200 import ops  # noqa: E402
from siss_b200.scheduler import SissDDPMScheduler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="default")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--filler", choices=["write", "read"], default="write",
                    help="what keeps the stream busy before each bracket: a 1 GB memset (leaves the 126 MB L2 full of DIRTY "
                         "lines, like K4b's 455 MB of output before K1oK2 in the real step) or a 1 GB read-only reduction "
                         "(leaves clean lines)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    dt = {"bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
    B, shape = args.batch, (args.batch, 3, 256, 256)
    D = 3 * 256 * 256
    s_in = 2 if dt == torch.bfloat16 else 4
    sched = SissDDPMScheduler()
    ac = sched.alphas_cumprod.to(dev)
    gamma, sigma = sched.gamma_sigma(dev)
    nsets = max(2, -(-(2 * 126 << 20) // (B * D * (4 * s_in + 12))))
    sets = []
    for i in range(nsets):
        g = torch.Generator(device=dev).manual_seed(i)
        x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dt)
        a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dt)
        nz = torch.randn(shape, device=dev, generator=g).to(dt)
        pred = torch.randn(shape, device=dev, generator=g)
        sets.append((x0, a0, nz, pred))
    t = torch.full((B,), 999, device=dev, dtype=torch.long)
    keep = (torch.rand(B, generator=torch.Generator().manual_seed(7)) > 0.5).to(torch.uint8).to(dev)
    go = 1.0 / B
    ev = {"k12": [], "k3": [], "copy12": [], "copy3": []}

    def bracket(name, fn):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = fn()
        e.record()
        ev[name].append((s, e))
        return r

    bytes12, bytes3 = 4 * s_in * B * D, (12 + 3 * s_in) * B * D
    c12s, c12d = torch.zeros(bytes12 // 2, dtype=torch.uint8, device=dev), torch.empty(bytes12 // 2, dtype=torch.uint8, device=dev)
    c3s, c3d = torch.zeros(bytes3 // 2, dtype=torch.uint8, device=dev), torch.empty(bytes3 // 2, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    # A filler that keeps the stream busy (~170 us) while the host enqueues the two bracketed launches: a bracket on an
    # IDLE stream would time the host's ctypes call, not the kernel (first version of this tool: 38 / 77 us instead of
    # 29 / 48). It also flushes L2, like the 2.3 GB of K4 traffic between two steps of bench.py.
    filler = torch.zeros(1 << 30, dtype=torch.uint8, device=dev)
    filler_i32 = filler.view(torch.int32)
    for i in range(args.iters + 5):
        x0, a0, nz, pred = sets[i % nsets]
        if args.filler == "write":
            filler.zero_()
        else:
            filler_i32.max()                    # reads 1 GB, writes one scalar: the L2 ends up full of CLEAN lines
        out = bracket("k12", lambda: ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5))
        x_mix, _, _, w_x, w_a = out
        bracket("k3", lambda: ops.wmse_fwd_bwd(pred, x_mix, x0, a0, t, gamma, sigma, w_x, w_a, go, go))
    torch.cuda.synchronize()
    for i in range(20):
        flush.zero_()
        bracket("copy12", lambda: c12d.copy_(c12s))
        flush.zero_()
        bracket("copy3", lambda: c3d.copy_(c3s))
    torch.cuda.synchronize()
    med = {k: statistics.median(a.elapsed_time(b) for a, b in v[2:]) * 1e3 for k, v in ev.items()}
    peak = 6447.5
    try:
        peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        pass
    knobs = {k: v for k, v in os.environ.items() if k.startswith("SISS_")}
    print(json.dumps({"tag": args.tag, "filler": args.filler, "knobs": knobs, "B": B, "dtype": args.dtype,
                      "k12_us": med["k12"], "k12_frac": bytes12 / med["k12"] / 1e3 / peak,
                      "k3_us": med["k3"], "k3_frac": bytes3 / med["k3"] / 1e3 / peak,
                      "copy12_us": med["copy12"], "copy3_us": med["copy3"]}), flush=True)


if __name__ == "__main__":
    main()
