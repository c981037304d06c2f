#!/usr/bin/env python
"""Per-kernel microbenchmark sweep (BASELINE.json configs[4]): achieved algorithmic GB/s of every
hot-path kernel over batch sizes / dtypes / gradient-buffer sizes. Writes JSON lines to stdout and
(optionally) a file. Each measurement rotates over enough distinct buffer sets that the working set
of consecutive iterations exceeds the 126 MB L2 (or says so when a single set is used)."""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

from siss_b200 import _lib, ops  # noqa: E402
from siss_b200.scheduler import SissDDPMScheduler  # noqa: E402

L2_BYTES = 126 << 20


def timeit(fn_list, iters=20, warmup=5):
    """Seconds per launch. The rotating launches are captured ONCE into a CUDA graph and the graph is replayed: the
    row then measures the kernel, not the host's launch path (round 1 timed eager launches, and every row below ~64
    samples reported the 15-35 us ctypes call instead of the kernel). Falls back to eager timing if capture fails."""
    n = len(fn_list)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(max(warmup, n)):
            fn_list[i % n]()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    per_graph = max(n, min(iters, 8))
    per_graph = (per_graph + n - 1) // n * n
    try:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(per_graph):
                fn_list[i % n]()
        reps = max(2, -(-iters // per_graph))
        graph.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            graph.replay()
        e.record()
        torch.cuda.synchronize()
        del graph
        return s.elapsed_time(e) / (reps * per_graph) * 1e-3
    except Exception:
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(iters):
            fn_list[i % n]()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters * 1e-3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--peak", type=float, default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak = args.peak
    if peak is None:
        try:
            peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
        except Exception:
            peak = 6650.0
    sched = SissDDPMScheduler()
    ac = sched.alphas_cumprod.to(dev)
    gamma, sigma = sched.gamma_sigma(dev)
    rows = []
    from siss_b200.rng import DeviceRng
    rng = DeviceRng(seed=42)

    def emit(**kw):
        kw["frac_of_peak"] = kw["gbs"] / peak
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    shapes = [("celeb", (3, 256, 256), [4, 64, 256, 1024, 4096] if not args.quick else [64, 256]),
              ("tshirt", (1, 28, 28), [32, 64, 4096] if not args.quick else [64]),
              ("sd", (4, 64, 64), [1, 16, 256] if not args.quick else [16])]
    for name, chw, batches in shapes:
        D = chw[0] * chw[1] * chw[2]
        for dt in (torch.bfloat16, torch.float32):
            s_in = 2 if dt == torch.bfloat16 else 4
            for B in batches:
                if B * D * 4 * 12 > 60e9:
                    continue
                per_set = B * D * (6 * s_in + 16)
                nsets = max(1, min(8, -(-2 * L2_BYTES // per_set)))
                sets = []
                for i in range(nsets):
                    g = torch.Generator(device=dev).manual_seed(i)
                    shape = (B,) + chw
                    x0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dt)
                    a0 = (torch.rand(shape, device=dev, generator=g) * 2 - 1).to(dt)
                    nz = torch.randn(shape, device=dev, generator=g).to(dt)
                    pred = torch.randn(shape, device=dev, generator=g)
                    pred2 = torch.randn(shape, device=dev, generator=g)      # No-IS: two DIFFERENT forward outputs
                    t = torch.full((B,), 999, device=dev, dtype=torch.long) if name != "tshirt" else \
                        torch.randint(0, 1000, (B,), device=dev, generator=g)
                    keep = (torch.rand(B, device=dev, generator=g) > 0.5).to(torch.uint8)
                    xt_x, xt_a = ops.add_noise_pair(x0, a0, nz, t, ac)
                    x_mix, _, _, w_x, w_a = ops.add_noise_mixture(x0, a0, nz, keep, t, ac, gamma, sigma, 0.5)
                    sets.append(dict(x0=x0, a0=a0, nz=nz, pred=pred, pred2=pred2, t=t, keep=keep, xt_x=xt_x, xt_a=xt_a, x_mix=x_mix,
                                     w_x=w_x, w_a=w_a))
                l2 = "rotating sets > L2" if nsets * per_set > L2_BYTES else f"L2-resident ({nsets * per_set >> 20} MiB)"
                N = B * D
                tests = {
                    "siss_add_noise_pair": (5 * s_in, [lambda z=z: ops.add_noise_pair(z["x0"], z["a0"], z["nz"], z["t"], ac) for z in sets]),
                    "siss_mixture_weights": (4 * s_in, [lambda z=z: ops.mixture_weights(z["xt_x"], z["xt_a"], z["x0"], z["a0"], z["keep"], z["t"], gamma, sigma, 0.5) for z in sets]),
                    "siss_add_noise_mixture": (4 * s_in, [lambda z=z: ops.add_noise_mixture(z["x0"], z["a0"], z["nz"], z["keep"], z["t"], ac, gamma, sigma, 0.5) for z in sets]),
                    # opt-in device RNG: eps generated in K1oK2's registers vs drawn by a separate
                    # launch and read back; torch.randn + K1oK2 is what the default path costs for the same work
                    "siss_add_noise_mixture_rng": (3 * s_in, [lambda z=z: ops.add_noise_mixture_rng(z["x0"], z["a0"], z["keep"], z["t"], ac, gamma, sigma, 0.5, 42, 7) for z in sets]),
                    "siss_randn": (s_in, [lambda z=z: rng.randn(z["nz"].shape, dt, dev, draw=7) for z in sets]),
                    "torch.randn": (s_in, [lambda z=z: torch.randn(z["nz"].shape, dtype=dt, device=dev) for z in sets]),
                    "torch.randn+siss_add_noise_mixture": (5 * s_in, [lambda z=z: ops.add_noise_mixture(z["x0"], z["a0"], torch.randn(z["nz"].shape, dtype=dt, device=dev), z["keep"], z["t"], ac, gamma, sigma, 0.5) for z in sets]),
                    "siss_wmse_fwd_bwd": (12 + 3 * s_in, [lambda z=z: ops.wmse_fwd_bwd(z["pred"], z["x_mix"], z["x0"], z["a0"], z["t"], gamma, sigma, z["w_x"], z["w_a"], 1 / 64, 1 / 64) for z in sets]),
                    "siss_wmse_fwd(api-compat)": (20 + 3 * s_in, [lambda z=z: ops.wmse_fwd(z["pred"], z["x_mix"], z["x0"], z["a0"], z["t"], gamma, sigma, z["w_x"], z["w_a"]) for z in sets]),
                    "siss_dual_mse_fwd_bwd": (16 + s_in, [lambda z=z: ops.dual_mse_fwd_bwd(z["pred"], z["pred2"], z["nz"], z["nz"], 1 / 64, 1 / 64) for z in sets]),
                }
                for k, (bpe, fns) in tests.items():
                    sec = timeit(fns)
                    emit(kernel=k, config=name, B=B, D=D, dtype=str(dt).split(".")[-1], us=sec * 1e6,
                         alg_bytes=bpe * N, gbs=bpe * N / sec / 1e9, samples_per_s=B / sec, l2=l2,
                         timing="CUDA-graph replay of the rotating launches")
                del sets
                torch.cuda.empty_cache()

    for P in ([100_000_000, 300_000_000, 900_000_000] if not args.quick else [113_673_220]):
        gx = torch.randn(P, device=dev) * 1e-3
        ga = torch.randn(P, device=dev) * 1e-3
        out = torch.empty_like(gx)
        sums = torch.zeros(3, dtype=torch.float64, device=dev)
        sec = timeit([lambda: ops.norm3(gx, ga, out=sums)])
        emit(kernel="siss_norm3", config="grad", P=P, us=sec * 1e6, alg_bytes=8 * P, gbs=8 * P / sec / 1e9, l2="buffers >> L2")
        sec = timeit([lambda: ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0, out=out)])
        emit(kernel="siss_combine", config="grad", P=P, us=sec * 1e6, alg_bytes=12 * P, gbs=12 * P / sec / 1e9, l2="buffers >> L2")

        def pair():
            ops.norm3(gx, ga, out=sums)
            ops.combine(gx, ga, sums, _lib.SISS_COMBINE_SCALING_NORM, 500.0, 1.0, out=out)
        sec = timeit([pair])
        emit(kernel="siss_norm3+siss_combine", config="grad", P=P, us=sec * 1e6, alg_bytes=20 * P, gbs=20 * P / sec / 1e9,
             l2="buffers >> L2 (K4b walks in reverse to reuse K4a's L2 tail)")
        del gx, ga, out
        torch.cuda.empty_cache()
    if args.out:
        Path(args.out).write_text("\n".join(json.dumps(r) for r in rows) + "\n")


if __name__ == "__main__":
    main()
