#!/bin/bash
# round 2, first GPU call: the whole -m gpu suite (timed per test), then a short bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r02_gpu.txt
nproc >> gpurun_out/r02_gpu.txt
( time timeout 1500 python -m pytest tests -q -m gpu --durations=25 -x ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -45 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_first.json 2> gpurun_out/r02_bench_first.err
tail -c 1500 gpurun_out/r02_bench_first.json
