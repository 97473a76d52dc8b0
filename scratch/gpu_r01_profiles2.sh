#!/bin/bash
# ncu captures of the kernels added late in round 1 (tracers, well-balanced split); run through gpurun.
set -x
NCU="ncu --clock-control none"
$NCU --set full --import-source on -k regex:tracer_recon -s 2 -c 1 -o gpurun_out/prof_tracer -f python bench.py --n 64 --avars 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tracer.log 2>&1
WB_N=40 $NCU --set full --import-source on -k regex:"eq_member|eq_solve|recon_kernel" -s 6 -c 3 -o gpurun_out/prof_wb3 -f python scratch/wb_prof.py > gpurun_out/ncu_wb3.log 2>&1
$NCU --metrics gpu__time_duration.sum -s 14 -c 21 --csv --log-file gpurun_out/launches_avars.csv python bench.py --n 64 --avars 1 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
WB_N=40 $NCU --metrics gpu__time_duration.sum -s 20 -c 15 --csv --log-file gpurun_out/launches_wb.csv python scratch/wb_prof.py > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_*.csv
