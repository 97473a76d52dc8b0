#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_properties.py -q -m gpu -x -k "wb or polytrope or atmosphere or recompute or equilibrium or heating or source" ) > gpurun_out/r02_pytest_e2.log 2>&1; tail -4 gpurun_out/r02_pytest_e2.log
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'])" || tail -3 ${1%.json}.err; }
timeout 600 python bench.py --kind polytrope2d --n 600 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/r02_e2_c2.json 2> gpurun_out/r02_e2_c2.err; show gpurun_out/r02_e2_c2.json "C2"
timeout 600 python bench.py --kind atmosphere --order 3 --n 64 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_e2_atm3.json 2> gpurun_out/r02_e2_atm3.err; show gpurun_out/r02_e2_atm3.json "atm o3"
timeout 600 python bench.py --kind atmosphere --order 4 --n 56 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_e2_c4.json 2> gpurun_out/r02_e2_c4.err; show gpurun_out/r02_e2_c4.json "C4"
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_launches_atm3_e2.csv python bench.py --kind atmosphere --order 3 --n 64 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
