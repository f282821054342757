#!/bin/bash
# full ncu capture of one steady-state rollout launch with extra environment: gpu_ncu_env.sh TAG "ENV=VAL ..." [bench args]
tag=$1; envs=$2; shift 2
out=gpurun_out; mkdir -p $out
env $envs ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_full -f python bench.py --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 --sweep 0 "$@" > $out/${tag}_ncu_full.log 2>&1
ls -la $out/${tag}_full.ncu-rep
