#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -q -m gpu -x ) > gpurun_out/r02_pytest_gpu_quick.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu_quick.log
timeout 600 python bench.py --kind atmosphere --order 4 --n 56 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_c4_o4b.json 2> gpurun_out/r02_bench_c4_o4.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c4_o4b.json').read().strip().splitlines()[-1]);print('C4 value',d['value']/1e9,'ms/step',d['ms_per_step'],d['roofline']['kernel_ms'])"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; python -c "
import json;d=json.loads(open('gpurun_out/r02_bench_c.json').read().strip().splitlines()[-1]);print('C3 value',d['value']/1e9,'ms/step',d['ms_per_step'],d['roofline']['kernel_ms'],'frac',d['roofline']['frac'],d['roofline']['stage']['frac'],'e2e',d['e2e']['value']/1e9,d['e2e']['rate_of_change']['value']/1e9,d['clocks'])"
