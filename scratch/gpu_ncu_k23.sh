#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"flux_kernel|update_kernel" -s 4 -c 2 -o gpurun_out/prof_k23 -f \
   python bench.py --n 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k23.log 2>&1
tail -2 gpurun_out/ncu_k23.log | cut -c 1-300
