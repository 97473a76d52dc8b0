#!/bin/bash
# K2 / K3 variants (frames read directly; K3 minimum blocks per SM) at 9.86 M tets
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
for v in base k2d k2d_k3b5 k2d_k3b4 base; do
  if [ $v = base ]; then unset ZFVM_LIB_PATH; else export ZFVM_LIB_PATH=$PWD/scratch/variants/libzfvm_$v.so; fi
  ZFVM_KNOB_DEFAULT_ONLY=$v timeout 600 python scratch/k1_knobs.py 118 3 2>&1 | grep -v "^setup" | tee -a gpurun_out/r02_k23_variants.log
done
