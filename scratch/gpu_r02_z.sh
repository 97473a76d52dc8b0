#!/bin/bash
# compute-sanitizer memcheck over the kernels added in round 2 (P1 lsq_weights, P2 stencil_search, E0 eq_decide, K2 with frames in registers)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py -q -m gpu -x -k "(device_stencil_search and (vortex2d_o3 or blast3d_o3 or open3d_o3 or six_stencils_o3)) or (bit_identical and (vortex_o3_hllc or blast_o3 or smooth3d_o4 or six)) or (recompute and 4-0.0002 and (polytrope_wb or atmosphere_wb_o4)) or (test_rate_of_change and (blast_o3 or vortex_o3_hllc or atmosphere_wb))" ) > gpurun_out/r02_sanitizer.log 2>&1; echo "exit $?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/r02_sanitizer.log | tail -8
