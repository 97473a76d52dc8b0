#!/bin/bash
# device-built weights: bit-identity test, the parity suite, bench with the create phases on stderr
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
( time timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "bit_identical" ) > gpurun_out/r02_pytest_weights.log 2>&1; tail -5 gpurun_out/r02_pytest_weights.log
( time ZFVM_VERBOSE=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_n118_devw.json 2> gpurun_out/r02_bench_n118_devw.err ); tail -1 gpurun_out/r02_bench_n118_devw.json | cut -c 1-2500; grep "zfvm" gpurun_out/r02_bench_n118_devw.err
( time timeout 1200 python -m pytest tests -q -m gpu -x ) > gpurun_out/r02_pytest_gpu2.log 2>&1; tail -5 gpurun_out/r02_pytest_gpu2.log
