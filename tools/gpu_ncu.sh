#!/bin/bash
# one full ncu capture of a steady-state rollout launch (+ launch list); outputs gpurun_out/<tag>_*
tag=${1:-p}; shift
out=gpurun_out; mkdir -p $out
python bench.py --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 --sweep 0 "$@" > $out/${tag}_plain.json 2> $out/${tag}_plain.err || { tail -5 $out/${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 --sweep 0 "$@" > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_full -f python bench.py --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 --sweep 0 "$@" > $out/${tag}_ncu_full.log 2>&1
ls -la $out | grep ${tag}_
