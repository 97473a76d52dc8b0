#!/bin/bash
# 2 GPUs: chunked host time step of decomposed runs (tests + the weak-scaling bench line with its e2e figure)
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests/test_gpu_distributed.py -q -x -k host_step ) > gpurun_out/r02_pytest_hoststep.log 2>&1; tail -15 gpurun_out/r02_pytest_hoststep.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_hoststep_weak_g2.json 2> gpurun_out/r02_hoststep_weak_g2.err
tail -1 gpurun_out/r02_hoststep_weak_g2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'roc %.4g' % d['e2e']['rate_of_change']['value'], d.get('parity_check'), 'setup', d['config']['setup_seconds'])" || tail -5 gpurun_out/r02_hoststep_weak_g2.err
