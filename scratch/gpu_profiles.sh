#!/bin/bash
# profiles for the round: launch list (64^3), full captures of K1 (118^3, the bench configuration) and K2/K3 (64^3)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/launches.csv \
   python bench.py --n 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"flux_face_kernel|update_kernel" -s 4 -c 2 -o gpurun_out/prof_k23 -f \
   python bench.py --n 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k23.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recon_tile -s 4 -c 1 -o gpurun_out/prof_k1_n118 -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k1.log 2>&1
tail -1 gpurun_out/ncu_k1.log | cut -c 1-200
ls -la gpurun_out/*.ncu-rep gpurun_out/launches.csv
