#!/bin/bash
# the contact path: "driving" distribution on a few tracks, 4096 envs
F="--steps 1000 --warmup 3000 --min-timed-steps 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0 --sweep 0 --config5 0 --steps-per-launch 200 --mode 1"
for tr in daytona martinsville nascar2 michigan; do
  python bench.py $F --track $tr 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['counters']
print('$tr'.ljust(14),'%.1f M'%(d['value']/1e6),'us/step %.2f'%(d['ms_per_step']*1e3),'contact steps',c['contact_steps'],'toi events',c['toi_events'])"
done
