#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_baseline_sizes.py -q -m gpu -x ) > gpurun_out/r02_pytest_gpu_quick.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu_quick.log
ZFVM_TILE_PROF=1 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_prof.log 2>&1
grep -a "tile prof\|phase\|default" gpurun_out/r02_k1_prof.log | head
ZFVM_TILE_PROF=1 ZFVM_TILE_L2_AHEAD=13824 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_prof_l2.log 2>&1
grep -a "tile prof\|phase\|default" gpurun_out/r02_k1_prof_l2.log | head
timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_knobs3.log 2>&1; cat gpurun_out/r02_k1_knobs3.log
