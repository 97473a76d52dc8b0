#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -q -m gpu --durations=10 -x ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu.log
timeout 300 python scratch/k1_l2.py 16 12 > gpurun_out/r02_k1_l2.log 2>&1; grep -a "tile prof\|K1" gpurun_out/r02_k1_l2.log | tail -4
ZFVM_VERBOSE=1 ZFVM_KNOB_GHOSTS_LAST=1 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_ghosts_last.log 2>&1; cat gpurun_out/r02_k1_ghosts_last.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; tail -c 1200 gpurun_out/r02_bench_b.json; tail -3 gpurun_out/r02_bench_b.err
