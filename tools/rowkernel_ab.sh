#!/bin/bash
mkdir -p gpurun_out
AB=gpurun_out/r2_rowkernel_ab.jsonl; : > $AB
python tools/rowkernel_ab.py --tag tma_static >> $AB 2>gpurun_out/ab.err
SISS_OVERSUB=2 python tools/rowkernel_ab.py --tag tma_over2 >> $AB 2>>gpurun_out/ab.err
SISS_NO_TMA=1 python tools/rowkernel_ab.py --tag ldg_static >> $AB 2>>gpurun_out/ab.err
for k in 2 4 8; do
  SISS_NO_TMA=1 SISS_LDG_OVERSUB=$k python tools/rowkernel_ab.py --tag ldg_over$k >> $AB 2>>gpurun_out/ab.err
done
cat $AB
