#!/bin/bash
# last call of round 2: GPU suite on the final library, the bench line with the set-up breakdown
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -q -m gpu ) > gpurun_out/r02_pytest_gpu_final3.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final3.log
ZFVM_VERBOSE=1 timeout 300 python bench.py > gpurun_out/r02_final3_c3.json 2> gpurun_out/r02_final3_c3.err
tail -1 gpurun_out/r02_final3_c3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']), 'e2e %.4g' % d['e2e']['value'], 'setup', d['config']['setup_seconds'], d['config']['setup_parts'], d['clocks'])" || tail -3 gpurun_out/r02_final3_c3.err
grep "zfvm" gpurun_out/r02_final3_c3.err | head -20
