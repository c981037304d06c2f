#!/usr/bin/env python
"""Phase-by-phase timing of the gradient-exchange kernels on N GPUs (launch with torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29555 \
        tools/exchange_probe.py [--params P] [--out gpurun_out/exchange_probe_wN.json]

For every kernel of the peer-memory (csrc/p2p.cu) and NVSwitch-multicast (csrc/nvls.cu) transports: device time per
launch with all ranks active (a symmetric-memory barrier precedes every launch, its cost is measured separately and
subtracted), the bytes that cross this GPU's NVLink ports in each direction, and the resulting GB/s per direction.
Then every complete schedule and the NCCL collectives (GradCombiner's start-up measurement, repeated with more
iterations). Rank 0 prints / writes one JSON document."""
import argparse
import ctypes
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--params", type=int, default=113_673_219)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    ap.add_argument("--schedules-only", action="store_true", help="skip the per-kernel phase timings")
    args = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    from siss_b200 import _lib
    from siss_b200.grad_combine import GradCombiner
    lib = _lib.load()
    holder = torch.nn.Parameter(torch.empty(args.params, device=dev))
    comb = GradCombiner([holder], transport="auto")
    pe = comb.peer
    res = {"world": world, "params": args.params, "total_padded": comb.total, "multicast": bool(pe and pe.has_multicast),
           "startup_tuning_ms": dict(comb.tuning),
           "startup_choice": {"full": comb.transport, "nccl_full": comb._nccl_full, "nccl_xpre": comb._nccl_xpre,
                              "peer_full": pe.algo if pe else None, "peer_xpre": pe.algo_xpre if pe else None,
                              "peer_three_stage": pe.algo3 if pe else None}}
    if pe is not None:
        P4, S4, N = 4 * comb.total, 4 * comb.shard_len, world
        g = torch.Generator(device=dev).manual_seed(rank)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        P = ctypes.c_void_p
        stats = comb.stats
        SN = _lib.SISS_COMBINE_SCALING_NORM

        def fill():
            pe.g_x.copy_(torch.randn(comb.total, device=dev, generator=g) * 1e-3)
            pe.g_a.copy_(torch.randn(comb.total, device=dev, generator=g) * 1e-3)

        def timeit(fn, iters=args.iters):
            fill()
            for _ in range(2):
                pe.h_x.barrier(channel=0); fn()
            torch.cuda.synchronize(); dist.barrier()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                pe.h_x.barrier(channel=0); fn()
            e.record()
            torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e) / iters], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        t_bar = timeit(lambda: None)
        phases = {"p2p_reduce_x_a": (lambda: pe._reduce(lib, stream, "p2p", 0), (N - 1) * S4 * 2, (N - 1) * S4 * 2),
                  "p2p_reduce_a": (lambda: pe._reduce(lib, stream, "p2p", 2), (N - 1) * S4, (N - 1) * S4),
                  "p2p_combine_allgather": (lambda: _lib.check(lib.siss_p2p_combine_allgather(
                      pe.shard_x.data_ptr(), pe.shard_a.data_ptr(), pe.scalars.data_ptr(), pe.ptrs_x, N, rank, comb.shard_len,
                      SN, 500.0, 1.0, 0, stats.data_ptr(), stream), "gather"), (N - 1) * S4, (N - 1) * S4)}
        if pe.has_multicast:
            phases.update({
                "nvls_reduce_x_a": (lambda: pe._reduce(lib, stream, "nvls", 0), 2 * P4, 2 * S4),
                "nvls_reduce_a": (lambda: pe._reduce(lib, stream, "nvls", 2), P4, S4),
                "nvls_combine_allgather": (lambda: _lib.check(lib.siss_nvls_combine_allgather(
                    pe.shard_x.data_ptr(), pe.shard_a.data_ptr(), pe.scalars.data_ptr(), P(pe.mc_x), N, rank, comb.shard_len,
                    SN, 500.0, 1.0, 0, stats.data_ptr(), stream), "gather"), S4, P4),
                "nvls_xcombine_bcast": (lambda: _lib.check(lib.siss_nvls_xcombine_bcast(
                    P(pe.mc_x), pe.shard_a.data_ptr(), pe.slots1.data_ptr(), pe.ptrs_s2, N, rank, comb.shard_len, 500.0, 0,
                    pe.ws.data_ptr(), stream), "xcombine"), P4 + S4, S4 + P4),
            })
        from siss_b200.p2p import CE_CHUNKS
        phases.update({
            "ce_reduce_x_a": (lambda: pe._reduce(lib, stream, "ce", 0), (N - 1) * S4 * 2, (N - 1) * S4 * 2),
            "ce_reduce_a": (lambda: pe._reduce(lib, stream, "ce", 2), (N - 1) * S4, (N - 1) * S4),
            "ce_combine_allgather": (lambda: _lib.check(lib.siss_ce_combine_allgather(
                pe.shard_x.data_ptr(), pe.shard_a.data_ptr(), pe.scalars.data_ptr(), pe.ptrs_x, N, rank, comb.shard_len,
                CE_CHUNKS, SN, 500.0, 1.0, 0, stats.data_ptr(), stream), "gather"), (N - 1) * S4, (N - 1) * S4)})
        phases["scale_finalize(local, 8 B/param HBM)"] = (lambda: _lib.check(lib.siss_scale_finalize(
            pe.g_x.data_ptr(), comb.total, pe.slots1.data_ptr(), pe.slots2.data_ptr(), N, 500.0, 1.0, 0, stats.data_ptr(),
            stream), "scale"), 0, 0)
        # make the slot arrays sane for the kernels that read them out of order here
        pe.combine(SN, 500.0, 1.0, False, stats, algo="p2p")
        if pe.has_multicast:
            pe.combine(SN, 500.0, 1.0, False, stats, algo="pipe")
        out = {"barrier_ms": t_bar}
        for name, (fn, b_out, b_in) in ({} if args.schedules_only else phases).items():
            t = timeit(fn) - t_bar
            out[name] = {"ms": t, "out_bytes": b_out, "in_bytes": b_in,
                         "gbs_out": b_out / t / 1e6 if b_out else None, "gbs_in": b_in / t / 1e6 if b_in else None}
        res["phases"] = out
        sched = {}
        for a in pe.available():
            sched[a] = timeit(lambda a=a: pe.combine(SN, 500.0, 1.0, False, stats, algo=a)) - t_bar
            if a in ("p2p", "nvls", "ce"):
                sched[a + "+xpre"] = timeit(lambda a=a: pe.combine(SN, 500.0, 1.0, False, stats, x_prereduced=True, algo=a)) - t_bar
        sched["nccl"] = timeit(lambda: comb._nccl_exchange(SN, 500.0, 1.0, False, False)) - t_bar
        sched["nccl+xpre"] = timeit(lambda: comb._nccl_exchange(SN, 500.0, 1.0, False, True)) - t_bar
        res["schedules_ms"] = sched
    if rank == 0:
        txt = json.dumps(res, indent=1)
        print(txt)
        if args.out:
            Path(args.out).parent.mkdir(parents=True, exist_ok=True)
            Path(args.out).write_text(txt)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
