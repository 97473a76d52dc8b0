#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -q -m gpu --durations=25 ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -30 gpurun_out/r02_pytest_gpu.log
timeout 600 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_knobs.log 2>&1
cat gpurun_out/r02_k1_knobs.log
