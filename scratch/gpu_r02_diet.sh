#!/bin/bash
# K1 instruction diet: A/B of variant builds on one box (scratch/build_variants.sh); usage: gpu_r02_diet.sh name[:lib] ...
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'K1frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 ${1%.json}.err; }
for v in "$@"; do
  name=${v%%:*}; lib=${v#*:}
  if [ "$lib" = "$v" ] || [ -z "$lib" ]; then unset ZFVM_LIB_PATH; else export ZFVM_LIB_PATH=$PWD/scratch/variants/libzfvm_$lib.so; fi
  timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02_diet_$name.json 2> gpurun_out/r02_diet_$name.err; show gpurun_out/r02_diet_$name.json $name
done
unset ZFVM_LIB_PATH
( timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "blast_o3 or vortex_o3_hllc or polynomial" 2>&1 | tail -2 )
