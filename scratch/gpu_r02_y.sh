#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
for v in base q2 base q2; do
  if [ $v = base ]; then unset ZFVM_LIB_PATH; else export ZFVM_LIB_PATH=$PWD/scratch/variants/libzfvm_$v.so; fi
  ZFVM_KNOB_DEFAULT_ONLY=$v timeout 600 python scratch/k1_knobs.py 118 3 2>&1 | grep -v "^setup" | tee -a gpurun_out/r02_k1_qstage.log
done
