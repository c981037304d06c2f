#!/usr/bin/env python
"""torchrun probe: time the pieces of the fused NVLink exchange (barrier, reduce-scatter+K4a kernel,
K4b+all-gather kernel) separately. Usage: torchrun --nproc-per-node N tools/p2p_probe.py [P]"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist

from siss_b200 import _lib, parallel
from siss_b200.p2p import PeerExchange

rank, world, local = parallel.init_from_env("nccl")
dev = torch.device("cuda", local)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 113_673_219
pad = 4 * world
Ptot = (P + pad - 1) // pad * pad
pe = PeerExchange(Ptot, dev)
pe.g_x.copy_(torch.randn(Ptot, device=dev) * 1e-3)
pe.g_a.copy_(torch.randn(Ptot, device=dev) * 1e-3)
lib = _lib.load()
stats = torch.zeros(5, device=dev)
stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def reduce_k():
    _lib.check(lib.siss_p2p_reduce_norm3(pe.ptrs_x, pe.ptrs_a, pe.ptrs_s, world, rank, pe.shard_len, pe.shard_x.data_ptr(),
                                         pe.shard_a.data_ptr(), pe.sums_local.data_ptr(), 0, pe.ws.data_ptr(), stream), "reduce")


def gather_k():
    _lib.check(lib.siss_p2p_combine_allgather(pe.shard_x.data_ptr(), pe.shard_a.data_ptr(), pe.scalars.data_ptr(), pe.ptrs_x,
                                              world, rank, pe.shard_len, 0, 500.0, 1.0, 0, stats.data_ptr(), stream), "gather")


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record(); torch.cuda.synchronize()
    t = torch.tensor([s.elapsed_time(e) / iters], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def with_barrier(fn):
    def g():
        pe.h_x.barrier(channel=0)
        fn()
    return g


res = {
    "barrier_only_ms": timeit(lambda: pe.h_x.barrier(channel=0)),
    "reduce_norm3_ms(+barrier)": timeit(with_barrier(reduce_k)),
    "combine_allgather_ms(+barrier)": timeit(with_barrier(gather_k)),
    "full_combine_ms": timeit(lambda: pe.combine(0, 500.0, 1.0, False, stats)),
}
S = pe.shard_len
res["reduce_inbound_GBs"] = (world - 1) * S * 8 / (res["reduce_norm3_ms(+barrier)"] - res["barrier_only_ms"]) / 1e6
res["gather_outbound_GBs"] = (world - 1) * S * 4 / (res["combine_allgather_ms(+barrier)"] - res["barrier_only_ms"]) / 1e6
# NCCL reference on the same buffers
shx = torch.empty(S, device=dev)
res["nccl_reduce_scatter_ms"] = timeit(lambda: dist.reduce_scatter_tensor(shx, pe.g_x))
res["nccl_all_gather_ms"] = timeit(lambda: dist.all_gather_into_tensor(pe.g_x, shx))
if rank == 0:
    print({k: round(v, 4) for k, v in res.items()}, "world", world, "P", P, file=sys.stderr)
dist.destroy_process_group()
