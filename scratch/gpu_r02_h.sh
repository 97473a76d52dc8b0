#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/repro_c1.py 158 100 > gpurun_out/r02_repro_c1.log 2>&1; tail -3 gpurun_out/r02_repro_c1.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python scratch/repro_c1.py 158 2 > gpurun_out/r02_repro_c1_sanitizer.log 2>&1; grep -v "^step" gpurun_out/r02_repro_c1_sanitizer.log | head -60
