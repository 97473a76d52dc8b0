#!/bin/bash
mkdir -p gpurun_out
for n in 14 16; do
ZFVM_TILE_EVICT=normal timeout 300 python scratch/k1_l2.py $n 12 > gpurun_out/r02_k1_l2_$n.log 2>&1; grep -a "tile prof\|K1" gpurun_out/r02_k1_l2_$n.log | tail -2
done
timeout 300 python scratch/k1_l2.py 16 12 > gpurun_out/r02_k1_l2_ef.log 2>&1; grep -a "tile prof\|K1" gpurun_out/r02_k1_l2_ef.log | tail -2
