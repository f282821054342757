#!/bin/bash
# the rollout kernel at 4096 envs (and one large batch) under both action distributions: one line per run
F="--steps 1000 --warmup 3000 --min-timed-steps 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0"
for cfg in "0 4096" "1 4096" "1 65536" "0 65536"; do
  set -- $cfg
  python bench.py $F --mode $1 --envs $2 --steps-per-launch 200 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mode',$1,'envs',$2,'value %.1f M'%(d['value']/1e6),'ms/step %.5f'%d['ms_per_step'], 'contact steps', d['counters']['contact_steps'])"
done
