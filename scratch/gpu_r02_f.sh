#!/bin/bash
mkdir -p gpurun_out
timeout 600 scratch/tma_ubench > gpurun_out/r02_tma_ubench.txt 2>&1; cat gpurun_out/r02_tma_ubench.txt
( time timeout 1800 python -m pytest tests -q -m gpu --durations=10 ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu.log
