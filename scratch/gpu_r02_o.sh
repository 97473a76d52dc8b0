#!/bin/bash
# sixteen-warp tile kernel for the small schemes: parity of the affected cases, then the 2D / 3D order-2 bench lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -q -m gpu -x -k "vortex or blast_o2 or polytrope or smooth3d_o2 or atmosphere_nowb or o2" ) > gpurun_out/r02_pytest_w16.log 2>&1; tail -4 gpurun_out/r02_pytest_w16.log
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']), d['config']['cells_per_gpu'], 'e2e %.4g' % (d['e2e']['value'] if d.get('e2e') else 0))" || tail -3 ${1%.json}.err; }
for ord in 3 2; do
timeout 600 python bench.py --kind vortex2d --n 1200 --order $ord --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_w16_2d_o$ord.json 2> gpurun_out/r02_w16_2d_o$ord.err; show gpurun_out/r02_w16_2d_o$ord.json "2d o$ord"
done
timeout 600 python bench.py --n 64 --order 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02_w16_3d_o2_n64.json 2> gpurun_out/r02_w16_3d_o2_n64.err; show gpurun_out/r02_w16_3d_o2_n64.json "3d o2 n64"
timeout 900 python bench.py --n 118 --order 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_w16_3d_o2_n118.json 2> gpurun_out/r02_w16_3d_o2_n118.err; show gpurun_out/r02_w16_3d_o2_n118.json "3d o2 n118"
timeout 600 python bench.py --kind vortex2d --n 158 --steps 200 --warmup 20 > gpurun_out/r02_w16_c1.json 2> gpurun_out/r02_w16_c1.err; show gpurun_out/r02_w16_c1.json "C1"
timeout 600 python bench.py --kind polytrope2d --n 600 --steps 20 --warmup 5 > gpurun_out/r02_w16_c2.json 2> gpurun_out/r02_w16_c2.err; show gpurun_out/r02_w16_c2.json "C2"
