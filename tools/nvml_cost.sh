#!/bin/bash
mkdir -p gpurun_out
for T in default nonvml; do
  if [ $T = nonvml ]; then export SISS_BENCH_NO_NVML=1; else unset SISS_BENCH_NO_NVML; fi
  python bench.py --no-e2e --no-extra-configs --no-cpu-baseline --no-copy-floor > gpurun_out/r2_nvml1_$T.json 2>/dev/null
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 50 --warmup 10 --no-e2e --no-extra-configs --no-cpu-baseline > gpurun_out/r2_nvml2_$T.json 2>/dev/null
  python - <<PY
import json
for f in ("gpurun_out/r2_nvml1_$T.json", "gpurun_out/r2_nvml2_$T.json"):
    d = json.loads(open(f).read().strip().splitlines()[-1])
    print("$T", d["n_gpus"], "ms/step", round(d["ms_per_step"], 4), "share", round(d["roofline"]["kernel_share_of_step"], 3), d["clocks"])
PY
done
