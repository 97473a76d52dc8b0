#!/bin/bash
# Round-1 closing measurements (run through gpurun from the repository root).
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
tail -c 600 gpurun_out/bench_main.json
# extra measurements: BASELINE config 4 (order 4, well-balanced), order 3 counterpart, config 2 scaled up, tracers
python bench.py --kind atmosphere --order 4 --n 56 --steps 5 --warmup 3 --cpu-n 12 > gpurun_out/bench_c4_o4.json 2> gpurun_out/bench_c4_o4.err
python bench.py --kind atmosphere --order 3 --n 64 --steps 5 --warmup 3 --cpu-n 16 > gpurun_out/bench_atm_o3.json 2> gpurun_out/bench_atm_o3.err
python bench.py --kind polytrope2d --order 3 --n 600 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
python bench.py --n 64 --avars 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_avars1.json 2> gpurun_out/bench_avars1.err
python bench.py --n 64 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n64.json 2> gpurun_out/bench_n64.err
for f in c4_o4 atm_o3 c2 avars1 n64; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g" % d["value"], d["roofline"]["kernel_ms"], "stage frac %.3f" % d["roofline"]["stage"]["frac"], "e2e %.4g" % d["e2e"]["value"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/bench_$f.err").read()[-800:])
PY
done
