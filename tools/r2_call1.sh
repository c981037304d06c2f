#!/bin/bash
# round-2 GPU call 1 (1 GPU): full GPU suite + row-kernel A/B (TMA ring vs LDG, static vs oversubscribed grids)
mkdir -p gpurun_out
rm -f gpurun_out/weights_pin.jsonl
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
AB=gpurun_out/r2_rowkernel_ab.jsonl; : > $AB
python tools/rowkernel_ab.py --tag tma_static >> $AB 2>gpurun_out/ab.err
SISS_OVERSUB=2 python tools/rowkernel_ab.py --tag tma_over2 >> $AB 2>>gpurun_out/ab.err
SISS_NO_TMA=1 python tools/rowkernel_ab.py --tag ldg_static >> $AB 2>>gpurun_out/ab.err
for k in 2 4 8 16; do
  SISS_NO_TMA=1 SISS_LDG_OVERSUB=$k python tools/rowkernel_ab.py --tag ldg_over$k >> $AB 2>>gpurun_out/ab.err
done
cat $AB
