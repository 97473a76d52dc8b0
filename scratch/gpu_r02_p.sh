#!/bin/bash
# ring geometry of the tile kernel (3D order 3, 9.86 M tets): variant builds of the library, K1 ms per launch
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm --format=csv,noheader
for v in base s4x4608 s5x3072 s4x3072 base; do
  if [ $v = base ]; then unset ZFVM_LIB_PATH; else export ZFVM_LIB_PATH=$PWD/scratch/variants/libzfvm_$v.so; fi
  ZFVM_KNOB_DEFAULT_ONLY=$v timeout 600 python scratch/k1_knobs.py 118 3 2>&1 | grep -v "^setup" | tee -a gpurun_out/r02_k1_ring.log
done
