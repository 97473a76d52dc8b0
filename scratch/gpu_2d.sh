#!/bin/bash
mkdir -p gpurun_out
for ord in 3 2 4; do
timeout 600 python bench.py --kind vortex2d --n 1200 --order $ord --steps 5 --warmup 3 --no-e2e > gpurun_out/b2d_$ord.log 2> gpurun_out/b2d_$ord.err
tail -1 gpurun_out/b2d_$ord.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('2d o$ord', 'value %.4g' % d['value'], d['roofline']['kernel_ms'], d['roofline']['algorithmic_bytes_per_cell'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']), d['config']['cells_per_gpu'])" || tail -3 gpurun_out/b2d_$ord.err
done
timeout 600 python bench.py --n 64 --order 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b3d_o2.log 2> gpurun_out/b3d_o2.err
tail -1 gpurun_out/b3d_o2.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('3d o2', 'value %.4g' % d['value'], d['roofline']['kernel_ms'], d['roofline']['algorithmic_bytes_per_cell'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']), d['config']['cells_per_gpu'])" || tail -3 gpurun_out/b3d_o2.err
exit 0
