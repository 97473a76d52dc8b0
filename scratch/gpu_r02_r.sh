#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "kernels_agree or source_paths or test_rate_of_change" ) > gpurun_out/r02_pytest_k1b.log 2>&1; tail -4 gpurun_out/r02_pytest_k1b.log
ZFVM_KNOB_DEFAULT_ONLY=tno_smem timeout 600 python scratch/k1_knobs.py 118 3 2>&1 | tee -a gpurun_out/r02_k1_ring.log
