#!/bin/bash
# K1: the per-segment L2 prefetch as one bulk instruction of lane 0 (ZFVM_TILE_L2_BULK=1) against a line per lane; A/B on one box
mkdir -p gpurun_out
show() { tail -1 $1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'K1frac %.3f' % d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" || tail -3 ${1%.json}.err; }
run() { name=$1; shift; env "$@" timeout 400 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r02_l2bulk_$name.json 2> gpurun_out/r02_l2bulk_$name.err; show gpurun_out/r02_l2bulk_$name.json $name; }
run base A=1
run bulk ZFVM_TILE_L2_BULK=1
run bulk2 ZFVM_TILE_L2_BULK=1 ZFVM_TILE_L2_AHEAD=9216
run base_again A=1
( ZFVM_TILE_L2_BULK=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "blast_o3 or vortex_o3_hllc" 2>&1 | tail -2 )
