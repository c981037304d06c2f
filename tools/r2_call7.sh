#!/bin/bash
mkdir -p gpurun_out
for T in pipe pipe_nvls nvls p2p nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 5 --transport $T --no-e2e --no-extra-configs --no-cpu-baseline > gpurun_out/r2_gap_$T.json 2> gpurun_out/r2_gap_$T.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_gap_$T.json").read().strip().splitlines()[-1])
print("$T", "ms/step", round(d["ms_per_step"], 4), "exchange bracket", round(d["comm"]["exchange_ms"], 4), "share", round(d["roofline"]["kernel_share_of_step"], 3), d["exchange_check"])
PY
done
