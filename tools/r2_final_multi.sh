#!/bin/bash
# final multi-GPU measurements of round 2: bash tools/r2_final_multi.sh <world> [probe sizes...]
mkdir -p gpurun_out
W=$1; shift
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $W --steps 50 --warmup 10 > gpurun_out/r2_final_bench_w$W.json 2> gpurun_out/r2_final_bench_w$W.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_final_bench_w$W.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_final_bench_w$W.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "exchange_check", "gpu_launches")}, d["comm"]["schedule"], {k: round(v, 3) for k, v in d["comm"]["tuning_ms"].items()}, "exchange", round(d["comm"]["exchange_ms"], 4), "share", round(d["roofline"]["kernel_share_of_step"], 3), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 3), (d.get("unlearn_steps") or {}).get("steps_per_s"), (d.get("unlearn_steps") or {}).get("transport"), d["clocks"])
PY
for P in "$@"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29555 tools/exchange_probe.py --schedules-only --iters 10 --params $P --out gpurun_out/r2_exchange_sweep_w${W}_P$P.json > gpurun_out/r2_probe_w$W.log 2>&1; echo "probe P=$P rc=$?"
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_exchange_sweep_w${W}_P$P.json"))
print($P, {k: round(v, 3) for k, v in d["schedules_ms"].items()})
PY
done
