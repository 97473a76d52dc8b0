#!/bin/bash
# final state of round 2: full GPU suite, smoke, bench lines of every BASELINE configuration
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader; nproc
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']), d['config']['cells_per_gpu'], 'e2e %.4g' % (d['e2e']['value'] if d.get('e2e') else 0), 'setup', d['config']['setup_seconds'], 'cpu', (d.get('cpu_baseline') or {}).get('value'))" || tail -3 ${1%.json}.err; }
timeout 900 python bench.py > gpurun_out/r02_final_c3.json 2> gpurun_out/r02_final_c3.err; show gpurun_out/r02_final_c3.json "C3 o3"
timeout 900 python bench.py --order 2 --no-cpu-baseline > gpurun_out/r02_final_c3_o2.json 2> gpurun_out/r02_final_c3_o2.err; show gpurun_out/r02_final_c3_o2.json "C3 o2"
timeout 600 python bench.py --kind vortex2d --n 158 --steps 200 --warmup 20 > gpurun_out/r02_final_c1.json 2> gpurun_out/r02_final_c1.err; show gpurun_out/r02_final_c1.json "C1"
timeout 600 python bench.py --kind polytrope2d --n 600 --steps 20 --warmup 5 > gpurun_out/r02_final_c2.json 2> gpurun_out/r02_final_c2.err; show gpurun_out/r02_final_c2.json "C2"
timeout 600 python bench.py --kind atmosphere --order 3 --n 64 --steps 10 --warmup 3 > gpurun_out/r02_final_atm3.json 2> gpurun_out/r02_final_atm3.err; show gpurun_out/r02_final_atm3.json "atm o3"
timeout 600 python bench.py --kind atmosphere --order 4 --n 56 --steps 10 --warmup 3 > gpurun_out/r02_final_c4.json 2> gpurun_out/r02_final_c4.err; show gpurun_out/r02_final_c4.json "C4"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_ref.json 2> gpurun_out/r02_final_ref.err; tail -1 gpurun_out/r02_final_ref.json | cut -c 1-600
