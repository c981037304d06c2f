#!/bin/bash
mkdir -p gpurun_out
W=2
timeout 900 python -m pytest "tests/test_exchange_fullsize_gpu.py::test_exchange_fullsize[$W]" -x -q > gpurun_out/r2_exch_test_w$W.log 2>&1; echo "rc=$?" >> gpurun_out/r2_exch_test_w$W.log; tail -5 gpurun_out/r2_exch_test_w$W.log
for C in 4 6; do
SISS_CE_CHUNKS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $W --master-addr 127.0.0.1 --master-port 29555 tools/exchange_probe.py --out gpurun_out/r2_exchange_probe_w${W}_c$C.json > gpurun_out/r2_probe_w$W.log 2>&1; echo "probe chunks=$C rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_exchange_probe_w${W}_c$C.json"))
print({k: round(v["ms"], 4) for k, v in d["phases"].items() if k.startswith("ce_")}, {k: round(v, 4) for k, v in d["schedules_ms"].items()})
PY
done
