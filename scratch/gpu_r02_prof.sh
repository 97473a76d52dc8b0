#!/bin/bash
# round-2 profiles: launch list (bench command), full capture of K1 at the bench size, K2/K3 at 64^3
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 24 --csv --log-file gpurun_out/r02_launches_n118.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recon_tile -s 4 -c 1 -o gpurun_out/r02_prof_k1_n118 -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_k1.log 2>&1
tail -1 gpurun_out/r02_ncu_k1.log | cut -c 1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"flux_face_kernel|update_kernel" -s 4 -c 2 -o gpurun_out/r02_prof_k23 -f \
   python bench.py --n 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_k23.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep gpurun_out/r02_launches_n118.csv
# bench lines of the other BASELINE configurations
timeout 600 python bench.py --kind atmosphere --order 4 --n 56 --steps 10 --warmup 3 > gpurun_out/r02_bench_c4_o4.json 2> gpurun_out/r02_bench_c4_o4.err; tail -c 600 gpurun_out/r02_bench_c4_o4.json
timeout 600 python bench.py --kind atmosphere --order 3 --n 64 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_atm_o3.json 2> gpurun_out/r02_bench_atm_o3.err; tail -c 400 gpurun_out/r02_bench_atm_o3.json
timeout 600 python bench.py --kind polytrope2d --n 600 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_c2.json 2> gpurun_out/r02_bench_c2.err; tail -c 400 gpurun_out/r02_bench_c2.json
timeout 600 python bench.py --kind vortex2d --n 158 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c1.json 2> gpurun_out/r02_bench_c1.err; tail -c 400 gpurun_out/r02_bench_c1.json
