#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:recon_tile -s 2 -c 1 -o gpurun_out/prof_tile -f \
   python bench.py --n 48 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tile.log 2>&1
tail -3 gpurun_out/ncu_tile.log
ls -la gpurun_out/*.ncu-rep
