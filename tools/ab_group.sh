#!/bin/bash
# group size (envs per CTA group) vs throughput at a given batch: NCG_GROUP_ENVS sweep, both launch shapes
F="--steps 1000 --warmup 3000 --min-timed-steps 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0 --sweep 0 --config5 0 --steps-per-launch 100"
for spec in "$@"; do
  E=${spec%%:*}; gs=${spec#*:}
  for g in $(echo $gs | tr ',' ' '); do
    for pw in 1 2; do
      NCG_GROUP_ENVS=$g NCG_PHYS_WARPS=$pw python bench.py $F --envs $E 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs $E group $g pw $pw  %.1f M'%(d['value']/1e6))"
    done
  done
done
