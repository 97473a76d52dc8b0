#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python scratch/diag_sizes.py all > gpurun_out/r02_diag_sizes.log 2>&1
tail -80 gpurun_out/r02_diag_sizes.log
ZFVM_TILE_PROF=1 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_prof.log 2>&1
grep -a "tile prof\|default" gpurun_out/r02_k1_prof.log | head
ZFVM_KNOB_GHOSTS_LAST=1 timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_ghosts_last.log 2>&1; cat gpurun_out/r02_k1_ghosts_last.log
