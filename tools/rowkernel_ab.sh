#!/bin/bash
# A/B of the two row kernels at the headline shape (tools/rowkernel_ab.py): TMA ring vs plain LDG, static vs
# oversubscribed grids, and other shapes of the shared-memory ring (CTAs per SM x stages)
mkdir -p gpurun_out
AB=gpurun_out/r2_rowkernel_ab.jsonl; : > $AB
python tools/rowkernel_ab.py --tag tma_static >> $AB 2>gpurun_out/ab.err
SISS_OVERSUB=2 python tools/rowkernel_ab.py --tag tma_over2 >> $AB 2>>gpurun_out/ab.err
SISS_NO_TMA=1 python tools/rowkernel_ab.py --tag ldg_static >> $AB 2>>gpurun_out/ab.err
for k in 2 4 8; do
  SISS_NO_TMA=1 SISS_LDG_OVERSUB=$k python tools/rowkernel_ab.py --tag ldg_over$k >> $AB 2>>gpurun_out/ab.err
done
SISS_K12_VARIANT=1 SISS_K3_VARIANT=1 python tools/rowkernel_ab.py --tag "ring_k12_2x9_k3_3x3" >> $AB 2>>gpurun_out/ab.err
SISS_K12_VARIANT=2 SISS_K3_VARIANT=2 python tools/rowkernel_ab.py --tag "ring_k12_3x4_k3_1x10" >> $AB 2>>gpurun_out/ab.err
SISS_K12_VARIANT=3 SISS_K3_VARIANT=3 python tools/rowkernel_ab.py --tag "ring_k12_1x18_k3_2x4" >> $AB 2>>gpurun_out/ab.err
cat $AB
