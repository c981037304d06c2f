#!/bin/bash
mkdir -p gpurun_out
for T in dur nvl; do
  if [ $T = dur ]; then export SISS_NCU_METRICS=gpu__time_duration.sum; else export SISS_NCU_METRICS=nvlrx__bytes.sum,nvltx__bytes.sum; fi
  export SISS_NCU_TAG=_$T
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29558 --no-python tools/ncu_rank0.sh tools/nvlink_capture.py > gpurun_out/r2_ncu_nvlink_$T.log 2>&1; echo "$T rc=$?"; grep -E "ERROR|ran " gpurun_out/r2_ncu_nvlink_$T.log | head -5
  head -c 1500 gpurun_out/r2_ncu_nvlink_w2_$T.csv
done
