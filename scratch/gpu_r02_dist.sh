#!/bin/bash
# usage: gpu_r02_dist.sh N [tests] [partitions]: strong scaling of the 118^3 bench mesh on N GPUs (SFC chunks / METIS k-way)
mkdir -p gpurun_out
N=${1:-2}
PARTS=${3:-"sfc metis"}
nvidia-smi --query-gpu=name --format=csv,noheader | head -$N | tr '\n' ';'; nproc
if [ "$2" = "tests" ]; then
( time timeout 900 python -m pytest tests/test_gpu_distributed.py -q -m gpu ) > gpurun_out/r02_pytest_dist.log 2>&1; tail -4 gpurun_out/r02_pytest_dist.log
fi
for part in $PARTS; do
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --scaling strong --partition $part --no-cpu-baseline > gpurun_out/r02_bench_strong_${part}_n118_g$N.json 2> gpurun_out/r02_bench_strong_${part}_n118_g$N.err ) 2>&1 | tail -3
tail -1 gpurun_out/r02_bench_strong_${part}_n118_g$N.json | cut -c 1-1500; tail -3 gpurun_out/r02_bench_strong_${part}_n118_g$N.err
done
