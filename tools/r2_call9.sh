#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest9.log; tail -6 gpurun_out/r2_pytest9.log
timeout 900 python bench.py > gpurun_out/r2_bench_w1.json 2> gpurun_out/r2_bench_w1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r2_bench_w1.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_w1.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")})
r = d["roofline"]; print(r["kernel"], r["frac"], r["kernel_share_of_step"], {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in r["kernels"].items()})
e = d["e2e"]; print("e2e", round(e["value"]), e["ms_per_step"], "minus stub", e["ms_per_step_minus_stub_unet"], e["breakdown_ms"], e.get("graph_variant"), e["device_rng_variant"]["ms_per_step"])
print("cpu", d["cpu_baseline"])
for k, v in (d.get("other_configs") or {}).items():
    print(k[:40], {kk: (vv if not isinstance(vv, dict) else {a: b for a, b in vv.items() if a in ("value", "ms_per_step", "ms_per_step_minus_stub_unet", "error", "cores")} ) for kk, vv in v.items() if kk in ("ms_per_step", "e2e", "cpu_baseline", "error")}, (v.get("e2e") or {}).get("graph_variant"))
PY
