#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
qb() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --n 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/qb_$name.log 2> gpurun_out/qb_$name.err
  tail -1 gpurun_out/qb_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'value %.4g' % d['value'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']))" || tail -3 gpurun_out/qb_$name.err
  grep "tile prof" gpurun_out/qb_$name.err
}
qb np
qb prof ZFVM_TILE_PROF=1
exit 0
