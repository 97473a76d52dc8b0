#!/bin/bash
# after the zfvm_create refactor: full GPU suite; final round-2 profiles (launch list of the bench command, ncu --set full of K1 at the
# bench size and of K2 / K3, durations of the two precompute kernels)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu -x ) > gpurun_out/r02_pytest_gpu_refactor.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_refactor.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_n118_final.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_launch_final.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recon_tile -s 4 -c 1 -o gpurun_out/r02_prof_k1_n118_final -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_k1_final.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"flux_face_kernel|update_kernel" -s 4 -c 2 -o gpurun_out/r02_prof_k23_final -f \
   python bench.py --n 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_k23_final.log 2>&1
ls -la gpurun_out/r02_*final*.ncu-rep gpurun_out/r02_launches_n118_final.csv
