#!/bin/bash
# last sanity pass of round 2 on 2 GPUs: the whole GPU suite (incl. the world-2 multi-GPU tests), smoke, a short N=2 bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_sanity_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_sanity_pytest.log; tail -4 gpurun_out/r2_sanity_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_sanity_bench_w2.json 2> gpurun_out/r2_sanity_bench_w2.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_sanity_bench_w2.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "exchange_check", "gpu_launches")}, d["comm"]["schedule"], "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), (d.get("unlearn_steps") or {}).get("steps_per_s"), d["clocks"])
PY
