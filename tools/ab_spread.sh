#!/bin/bash
# A/B on the GPU box: one physics warp per CTA vs the spread shape (four physics warps of eight lanes), both action distributions
out=gpurun_out
F="--steps 1000 --warmup 3000 --min-timed-steps 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0"
for pw in 1 4; do
  for mode in 0 1; do
    for envs in 4096; do
      NCG_PHYS_WARPS=$pw python bench.py $F --mode $mode --envs $envs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pw',$pw,'mode',$mode,'envs',$envs,'value %.1f M'%(d['value']/1e6),'ms/step %.5f'%d['ms_per_step'], d['counters']['contact_steps'])"
    done
  done
done
