#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --n 64 --order 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b3d_o2.log 2> gpurun_out/b3d_o2.err
tail -1 gpurun_out/b3d_o2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('3d o2', 'value %.4g' % d['value'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']))" || tail -3 gpurun_out/b3d_o2.err
ZFVM_TILE_PROF=1 timeout 600 python bench.py --n 64 --order 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>&1 | grep "tile prof"
exit 0
