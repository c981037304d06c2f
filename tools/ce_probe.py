#!/usr/bin/env python
"""Copy-engine (DMA) peer copies over NVLink, all ranks active at once (torchrun, one rank per GPU).

The fused exchange kernels move their bytes with SM loads / stores and reach ~530-600 GB/s per direction; a DMA copy
of a peer buffer (cudaMemcpyAsync on a mapped peer pointer) is reported at ~770 GB/s. This probe measures, with every
rank doing the same thing simultaneously on symmetric-memory buffers of the gradient size:
  pull      every rank copies ITS shard of every peer's buffer into local staging (one stream per peer)
  push      every rank copies its shard into every peer's buffer
  pull+push both at once (the two directions of each link)
and prints GB/s per direction per GPU (bytes that cross one GPU's ports in that direction / time)."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
    import torch.distributed._symmetric_memory as symm_mem
    P = 113_673_224 // (4 * world) * (4 * world)
    S = P // world
    buf = symm_mem.empty(P, dtype=torch.float32, device=dev)
    buf.fill_(float(rank))
    h = symm_mem.rendezvous(buf, dist.group.WORLD)
    peers = [h.get_buffer(r, (P,), torch.float32) for r in range(world)]
    staging = torch.empty(world, S, device=dev)
    mine = torch.full((S,), 7.0, device=dev)
    streams = [torch.cuda.Stream() for _ in range(world)]
    order = [(rank + k) % world for k in range(1, world)]          # rotate so that no peer is hit by everyone at once
    res = {"world": world, "shard_MB": S * 4 / 1e6}

    def pull(chunks=1):
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        c = S // chunks
        for r in order:
            st = streams[r]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                for k in range(chunks):
                    lo = k * c
                    hi = S if k == chunks - 1 else lo + c
                    staging[r, lo:hi].copy_(peers[r][rank * S + lo:rank * S + hi], non_blocking=True)
        for r in order:
            cur.wait_stream(streams[r])

    def push(chunks=1):
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        c = S // chunks
        for r in order:
            st = streams[r]
            st.wait_event(ev)
            with torch.cuda.stream(st):
                for k in range(chunks):
                    lo = k * c
                    hi = S if k == chunks - 1 else lo + c
                    peers[r][rank * S + lo:rank * S + hi].copy_(mine[lo:hi], non_blocking=True)
        for r in order:
            cur.wait_stream(streams[r])

    pstreams = [torch.cuda.Stream() for _ in range(world)]

    def both():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event(); ev.record(cur)
        for r in order:
            streams[r].wait_event(ev)
            pstreams[r].wait_event(ev)
            with torch.cuda.stream(streams[r]):
                staging[r].copy_(peers[r][rank * S:(rank + 1) * S], non_blocking=True)
            with torch.cuda.stream(pstreams[r]):
                peers[r][rank * S:(rank + 1) * S].copy_(mine, non_blocking=True)
        for r in order:
            cur.wait_stream(streams[r]); cur.wait_stream(pstreams[r])

    def timeit(fn, iters=10):
        for _ in range(2):
            h.barrier(channel=0); fn()
        torch.cuda.synchronize(); dist.barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            h.barrier(channel=0); fn()
        e.record()
        torch.cuda.synchronize()
        t = torch.tensor([s.elapsed_time(e) / iters], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    t_bar = timeit(lambda: None)
    nbytes = (world - 1) * S * 4
    for name, fn in (("pull", pull), ("pull_4chunks", lambda: pull(4)), ("push", push), ("push_4chunks", lambda: push(4)),
                     ("pull+push", both)):
        t = timeit(fn) - t_bar
        res[name] = {"ms": t, "GBps_per_direction": nbytes / t / 1e6}
    # correctness of the plumbing: staged rows hold the peers' rank ids, and every peer wrote 7.0 into its shard of ours
    pull(); torch.cuda.synchronize(); dist.barrier()
    ok = all(float(staging[r].min()) == float(staging[r].max()) for r in order)
    res["plumbing_ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(res, indent=1))
        out = ROOT / "gpurun_out"
        out.mkdir(exist_ok=True)
        (out / f"r2_ce_probe_w{world}.json").write_text(json.dumps(res, indent=1))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
