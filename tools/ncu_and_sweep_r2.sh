#!/bin/bash
mkdir -p gpurun_out
bash tools/ncu_r2.sh 2>&1 | tail -6
timeout 1500 python tools/microbench.py --out gpurun_out/sweep_r2.jsonl > gpurun_out/sweep_r2.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/sweep_r2.log | cut -c1-300
