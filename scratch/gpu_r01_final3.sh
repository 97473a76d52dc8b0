#!/bin/bash
# Round-1 closing bench lines after the well-balanced tile path (through gpurun, repository root).
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_main.json 2> gpurun_out/bench_main.err
python bench.py --kind atmosphere --order 3 --n 64 --steps 5 --warmup 3 --cpu-n 16 > gpurun_out/bench_atm_o3.json 2> gpurun_out/bench_atm_o3.err
python bench.py --kind polytrope2d --order 3 --n 600 --steps 5 --warmup 3 --cpu-n 20 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
for f in main atm_o3 c2; do python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", "%.4g" % d["value"], {k: (round(v, 3) if v else v) for k, v in d["roofline"]["kernel_ms"].items()}, "K1 frac %.3f stage frac %.3f" % (d["roofline"]["frac"], d["roofline"]["stage"]["frac"]), "e2e %.4g" % d["e2e"]["value"], "cpu %.4g" % (d["cpu_baseline"] or {"value": 0})["value"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/bench_$f.err").read()[-800:])
PY
done
