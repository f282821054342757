#!/bin/bash
# one physics warp per CTA vs the pair shape over batch sizes (single-car envs on daytona): where the launch rule should switch
F="--steps 1000 --warmup 3000 --min-timed-steps 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0 --sweep 0 --config5 0 --steps-per-launch 100"
for E in 12288 16384 20480 24576 32768 40960 49152 57344 65536 98304 131072; do
  for pw in 1 2; do
    NCG_PHYS_WARPS=$pw python bench.py $F --envs $E 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('envs $E pw $pw  %.1f M'%(d['value']/1e6))"
  done
done
