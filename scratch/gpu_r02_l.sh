#!/bin/bash
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -q -m gpu --durations=5 ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -8 gpurun_out/r02_pytest_gpu.log
timeout 300 python scratch/k1_knobs.py 118 3 > gpurun_out/r02_k1_knobs4.log 2>&1; cat gpurun_out/r02_k1_knobs4.log
timeout 600 python bench.py --kind vortex2d --n 158 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c1_graph.json 2> gpurun_out/r02_bench_c1.err; tail -c 300 gpurun_out/r02_bench_c1_graph.json | head -c 100; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c1_graph.json').read().strip().splitlines()[-1]);print('C1 value',d['value']/1e9,'ms/step',d['ms_per_step'])"
for cfg in "atmosphere 4 56 c4" "atmosphere 3 64 atm3" "polytrope2d 3 600 c2"; do set -- $cfg
ZFVM_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 30 --csv --log-file gpurun_out/r02_launches_$4.csv python bench.py --kind $1 --order $2 --n $3 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_$4.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_$4.csv')) if len(r)>10 and r[0].isdigit()]
for r in rows[:14]: print('$4', r[4][:70].ljust(70), r[-1])
PY
done
