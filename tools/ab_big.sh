#!/bin/bash
# large-batch workloads of the rollout kernel (random policy): one line per run (extra env vars apply to all)
F="--steps 1000 --warmup 3000 --min-timed-steps 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0 --sweep 0 --config5 0"
run() { python bench.py $F "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*'.ljust(60),'value %.1f M'%(d['value']/1e6),'us/step %.2f'%(d['ms_per_step']*1e3))"; }
run --envs 16384 --steps-per-launch 200
run --envs 65536 --steps-per-launch 100
run --envs 65536 --track all --steps-per-launch 100
run --envs 8192 --cars 10 --track talladega --steps-per-launch 100
