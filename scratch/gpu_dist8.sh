#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
ZFVM_BENCH_N=${2:-64} timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.log | cut -c 1-1200; tail -3 gpurun_out/bench_n$N.err
