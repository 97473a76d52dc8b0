#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi2.txt
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/pytest_dist.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dist.log
tail -5 gpurun_out/pytest_dist.log
ZFVM_BENCH_N=${1:-64} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.log 2> gpurun_out/bench_n2.err
tail -1 gpurun_out/bench_n2.log; tail -3 gpurun_out/bench_n2.err
