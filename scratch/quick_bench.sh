#!/bin/bash
# usage: quick_bench.sh <n> [extra bench args]; prints value, kernel ms, roofline frac
n=$1; shift
python bench.py --n $n --steps 5 --warmup 3 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value %.4g' % d['value'], d['roofline']['kernel_ms'], 'K1frac %.3f stagefrac %.3f' % (d['roofline']['frac'], d['roofline']['stage']['frac']))"
