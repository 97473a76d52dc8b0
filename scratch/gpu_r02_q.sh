#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
( time timeout 600 python -m pytest tests/test_gpu_extensions.py -q -m gpu -k "recompute" ) > gpurun_out/r02_pytest_rc.log 2>&1; tail -30 gpurun_out/r02_pytest_rc.log
