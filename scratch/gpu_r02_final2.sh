#!/bin/bash
# end of round 2, final code (bulk L2 prefetch in K1, chunked host routes for decomposed runs): full GPU suite, smoke, the
# driver's bench line and reference arm, launch list + ncu --set full of K1 for the committed roofline.traffic
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader
( time timeout 1500 python -m pytest tests -q -m gpu ) > gpurun_out/r02_pytest_gpu_final2.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r02_final2_c3.json 2> gpurun_out/r02_final2_c3.err
tail -1 gpurun_out/r02_final2_c3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('C3', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']), 'e2e %.4g' % d['e2e']['value'], 'setup', d['config']['setup_seconds'], d['config']['setup_parts'], 'cpu', d['cpu_baseline']['value'], d['clocks'], 'launches', d['gpu_launches'])" || tail -3 gpurun_out/r02_final2_c3.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final2_ref.json 2> gpurun_out/r02_final2_ref.err; tail -1 gpurun_out/r02_final2_ref.json | cut -c 1-400
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_n118_final2.csv \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_launch_final2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:recon_tile -s 4 -c 1 -o gpurun_out/r02_prof_k1_n118_final2 -f \
   python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r02_ncu_k1_final2.log 2>&1
ls -la gpurun_out/r02_*final2*
