#!/bin/bash
mkdir -p gpurun_out
W=${1:-2}
timeout 600 python -m pytest "tests/test_exchange_fullsize_gpu.py::test_exchange_fullsize[$W]" "tests/test_distributed_gpu.py::test_multi_gpu_step_equals_single_gpu[$W]" -x -q > gpurun_out/r2_overlap_test_w$W.log 2>&1; echo "rc=$?" >> gpurun_out/r2_overlap_test_w$W.log; tail -15 gpurun_out/r2_overlap_test_w$W.log
for T in overlap nooverlap; do
  if [ $T = nooverlap ]; then export SISS_NO_OVERLAP=1; else unset SISS_NO_OVERLAP; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $W --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2_overlap_bench_w${W}_$T.json 2> gpurun_out/r2_overlap_bench_w${W}_$T.err; echo "bench $T rc=$?"; tail -c 200 gpurun_out/r2_overlap_bench_w${W}_$T.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/r2_overlap_bench_w${W}_$T.json").read().strip().splitlines()[-1])
print("$T", "value ms", round(d["ms_per_step"], 4), d["exchange_check"], "e2e ms", round(d["e2e"]["ms_per_step"], 4), {k: round(v, 3) for k, v in d["e2e"]["breakdown_ms"].items()}, "unlearn", d["unlearn_steps"].get("steps_per_s"), d["unlearn_steps"].get("ms_per_step"), d["unlearn_steps"].get("error"))
PY
done
