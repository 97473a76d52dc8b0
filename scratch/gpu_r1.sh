#!/bin/bash
# smoke -> parity tests -> quick benches of the reconstruction kernel variants (64^3 x 6 tets)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -2 gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
qb() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --n 64 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/qb_$name.log 2> gpurun_out/qb_$name.err
  tail -1 gpurun_out/qb_$name.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$name', 'value %.4g' % d['value'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']))" || tail -3 gpurun_out/qb_$name.err
}
qb tile4 ZFVM_TILE_WARPS=4
qb tile3 ZFVM_TILE_WARPS=3
qb tile2 ZFVM_TILE_WARPS=2
qb stream ZFVM_RECON=stream
