#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scratch/diag_sizes.py all > gpurun_out/r02_diag_sizes2.log 2>&1
grep -a "max err\|tile vs\|step 19\|step 3:\|bad True" gpurun_out/r02_diag_sizes2.log | head -30
( time timeout 1800 python -m pytest tests -q -m gpu --durations=10 ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -15 gpurun_out/r02_pytest_gpu.log
ZFVM_KNOB_GHOSTS_LAST=1 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_ghosts_last.log 2>&1; cat gpurun_out/r02_k1_ghosts_last.log
