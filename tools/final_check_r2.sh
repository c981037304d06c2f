#!/bin/bash
# the driver's round-end sequence on one fresh box: GPU tests, smoke(), bench.py --impl reference, bench.py
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_final_pytest.log; tail -4 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_final_ref.json 2>/dev/null; echo "ref rc=$?"
timeout 1200 python bench.py > gpurun_out/r2_final_bench_w1.json 2> gpurun_out/r2_final_bench_w1.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2_final_bench_w1.err
python - <<PY
import json
r = json.loads(open("gpurun_out/r2_final_ref.json").read().strip().splitlines()[-1])
d = json.loads(open("gpurun_out/r2_final_bench_w1.json").read().strip().splitlines()[-1])
print("ref", round(r["value"], 2), r["cpu_baseline"]["cores"], "same config:", r["config"] == d["config"])
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d["clocks"])
ro = d["roofline"]; print(ro["kernel"], ro["frac"], ro["kernel_share_of_step"], {k: (round(v["ms"], 4), round(v["frac"], 3)) for k, v in ro["kernels"].items()})
e = d["e2e"]; print("e2e", round(e["value"]), round(e["ms_per_step"], 4), "minus stub", round(e["ms_per_step_minus_stub_unet"], 4), (e.get("graph_variant") or {}).get("ms_per_step"))
print("e2e ratio vs ref", e["value"] / r["value"])
for k, v in (d.get("other_configs") or {}).items():
    ee = v.get("e2e") or {}
    print(k[:32], v.get("ms_per_step"), ee.get("ms_per_step"), (ee.get("graph_variant") or {}).get("ms_per_step"), ee.get("error"), (v.get("cpu_baseline") or {}).get("ms_per_step"))
PY
