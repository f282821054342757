#!/bin/bash
# Run on the GPU box (via gpurun): GPU tests, the default bench line, the reference arm, the ncu launch list of a short
# bench command and one full capture of a steady-state rollout launch.  Outputs land in gpurun_out/<tag>_*.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
python bench.py --impl reference > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
SHORT="--steps 1000 --warmup 1000 --e2e-steps 20 --cpu-steps 200 --sweep 0"    # 3 warm-up launches + 1 timed launch of 1000 steps, then 25 single steps
python bench.py $SHORT > $out/${tag}_plain_short.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_launches.csv \
    python bench.py $SHORT > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 3 --launch-count 1 \
    -o $out/${tag}_full -f python bench.py $SHORT > $out/${tag}_ncu_full.log 2>&1
# the single-step (T=1) kernel of the e2e path: launches 8.. are NascarVectorEnv.step
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 14 --launch-count 1 \
    -o $out/${tag}_full_t1 -f python bench.py $SHORT > $out/${tag}_ncu_full_t1.log 2>&1
# a large batch (65536 envs: ray queue, pairs of env groups per CTA), one steady-state launch of 100 steps
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_full_big -f python bench.py --envs 65536 --steps-per-launch 100 --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 --sweep 0 > $out/${tag}_ncu_full_big.log 2>&1
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench.json; cat $out/${tag}_bench_ref.json
