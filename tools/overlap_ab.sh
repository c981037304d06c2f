#!/bin/bash
mkdir -p gpurun_out
W=${1:-8}
timeout 600 python -m pytest "tests/test_distributed_gpu.py::test_multi_gpu_step_equals_single_gpu[$W]" -x -q > gpurun_out/r2_regions_test_w$W.log 2>&1; echo "rc=$?" >> gpurun_out/r2_regions_test_w$W.log; tail -8 gpurun_out/r2_regions_test_w$W.log
export SISS_OVERLAP_REGIONS=4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $W --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_regions_bench_w$W.json 2> gpurun_out/r2_regions_bench_w$W.err; echo "bench rc=$?"; tail -c 200 gpurun_out/r2_regions_bench_w$W.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_regions_bench_w$W.json").read().strip().splitlines()[-1])
print("regions=4", "value ms", round(d["ms_per_step"], 4), d["exchange_check"], "e2e ms", round(d["e2e"]["ms_per_step"], 4), {k: round(v, 3) for k, v in d["e2e"]["breakdown_ms"].items()}, "unlearn", d["unlearn_steps"].get("steps_per_s"), d["unlearn_steps"].get("ms_per_step"), d["unlearn_steps"].get("error"))
PY
