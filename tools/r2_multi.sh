#!/bin/bash
# round-2 multi-GPU call: usage  bash tools/r2_multi.sh "<world sizes>"   e.g. "2" or "4 8"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for W in $1; do
  echo "=== world $W: full-size exchange parity"
  timeout 900 python -m pytest "tests/test_exchange_fullsize_gpu.py::test_exchange_fullsize[$W]" -x -q > gpurun_out/r2_exch_test_w$W.log 2>&1; echo "rc=$?" >> gpurun_out/r2_exch_test_w$W.log; tail -12 gpurun_out/r2_exch_test_w$W.log
  if [ "$W" != "4" ]; then
  echo "=== world $W: step-level N-rank == 1-rank"
  timeout 900 python -m pytest "tests/test_distributed_gpu.py::test_multi_gpu_step_equals_single_gpu[$W]" -x -q > gpurun_out/r2_dist_test_w$W.log 2>&1; echo "rc=$?" >> gpurun_out/r2_dist_test_w$W.log; tail -12 gpurun_out/r2_dist_test_w$W.log
  fi
  echo "=== world $W: exchange probe"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29555 tools/exchange_probe.py --iters 10 --out gpurun_out/r2_exchange_probe_w$W.json > gpurun_out/r2_probe_w$W.log 2>&1; echo "probe rc=$?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_exchange_probe_w$W.json"))
    print({k: (round(v["ms"], 4), round(v["gbs_out"] or 0), round(v["gbs_in"] or 0)) for k, v in d["phases"].items() if isinstance(v, dict)})
    print({k: round(v, 4) for k, v in d["schedules_ms"].items()}, d["startup_choice"])
except Exception as e:
    print("probe parse failed", e)
PY
  if [ "$W" = "8" ]; then
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29557 tools/ce_probe.py 2>&1 | grep -E "GBps|ms|plumbing" | tr -d '\n'; echo
  fi
  echo "=== world $W: bench"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $W --steps 20 --warmup 5 > gpurun_out/r2_bench_w$W.json 2> gpurun_out/r2_bench_w$W.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_bench_w$W.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_w$W.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "exchange_check", "gpu_launches")}, d["comm"]["schedule"], d["comm"]["tuning_ms"], "e2e", d["e2e"]["value"], (d.get("unlearn_steps") or {}).get("steps_per_s"), (d.get("unlearn_steps") or {}).get("transport"))
except Exception as e:
    print("bench parse failed", e)
PY
done
