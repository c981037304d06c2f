#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log; tail -6 gpurun_out/r2_pytest3.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 tools/ce_probe.py 2>&1 | tail -40
python tools/latency_probe.py 2>&1 | tail -8
