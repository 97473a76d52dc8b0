#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
nproc
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_full_n$N.log 2> gpurun_out/bench_full_n$N.err ) 2>&1 | tail -3
tail -1 gpurun_out/bench_full_n$N.log | cut -c 1-2500; tail -3 gpurun_out/bench_full_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c 1-300
