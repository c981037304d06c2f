#!/bin/bash
mkdir -p gpurun_out
nproc; cat /proc/cpuinfo | grep -c processor
W=4
run() { tag=$1; shift; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $W --warmup 5 --no-e2e --no-extra-configs --no-cpu-baseline "$@" > gpurun_out/r2_gap4_$tag.json 2> gpurun_out/r2_gap4_$tag.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_gap4_$tag.json").read().strip().splitlines()[-1])
print("$tag", d["comm"]["schedule"], "ms/step", round(d["ms_per_step"], 4), "exchange bracket", round(d["comm"]["exchange_ms"], 4), "share", round(d["roofline"]["kernel_share_of_step"], 3), "host_enqueue_ms", round(d["roofline"]["host_enqueue_ms_per_step"], 4))
PY
}
run default --steps 20
run steps100 --steps 100
SISS_BENCH_NO_NVML=1 run nonvml --steps 20
run p2p --steps 20 --transport p2p
run nccl --steps 20 --transport nccl
