#!/bin/bash
# 4 GPUs: the driver's weak-scaling command with the chunked host routes of decomposed runs (e2e figure, parity_check)
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_hoststep_weak_g4.json 2> gpurun_out/r02_hoststep_weak_g4.err
tail -1 gpurun_out/r02_hoststep_weak_g4.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], 'roc %.4g' % d['e2e']['rate_of_change']['value'], d.get('parity_check'), 'setup', d['config']['setup_seconds'])" || tail -5 gpurun_out/r02_hoststep_weak_g4.err
