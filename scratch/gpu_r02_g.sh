#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -12 gpurun_out/r02_pytest_gpu.log
ZFVM_TILE_PROF=1 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_prof.log 2>&1
grep -a "tile prof\|phase\|default" gpurun_out/r02_k1_prof.log | head
timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_knobs2.log 2>&1; cat gpurun_out/r02_k1_knobs2.log
timeout 900 python scratch/diag_sizes.py o2s > gpurun_out/r02_diag_o2s.log 2>&1; grep -a "step\|max err" gpurun_out/r02_diag_o2s.log | head -60
