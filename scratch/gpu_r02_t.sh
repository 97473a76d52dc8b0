#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
( time ZFVM_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "device_stencil_search" ) > gpurun_out/r02_pytest_stsearch.log 2>&1; grep "left to the host\|passed\|failed\|Error" gpurun_out/r02_pytest_stsearch.log | tail -20
( time ZFVM_VERBOSE=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n118_devst.json 2> gpurun_out/r02_bench_n118_devst.err ); tail -1 gpurun_out/r02_bench_n118_devst.json | cut -c 1-900; grep "zfvm" gpurun_out/r02_bench_n118_devst.err
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r02_pytest_gpu4.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu4.log
