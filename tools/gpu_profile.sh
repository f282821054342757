#!/bin/bash
# Run on the GPU box (via gpurun): GPU tests, the bench line as the driver runs it, the reference arm, the ncu launch list of
# a short bench command and full captures of steady-state launches.  Outputs land in gpurun_out/<tag>_*.
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,driver_version --format=csv > $out/${tag}_gpu.txt 2>&1
nproc >> $out/${tag}_gpu.txt
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err || { tail -5 $out/${tag}_bench.err; exit 1; }
# 3 warm-up launches + 1 timed launch of 1000 steps, then 25 single steps through NascarVectorEnv.step
SHORT="--steps 1000 --warmup 1000 --min-timed-steps 1000 --min-warmup 1000 --e2e-steps 20 --extras 0 --cpu-baseline 0"
python bench.py $SHORT > $out/${tag}_plain_short.json 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/${tag}_launches.csv \
    python bench.py $SHORT > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 3 --launch-count 1 \
    -o $out/${tag}_full -f python bench.py $SHORT > $out/${tag}_ncu_full.log 2>&1
# the single-step (T=1) kernel of the e2e path: launches 8.. are NascarVectorEnv.step
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 14 --launch-count 1 \
    -o $out/${tag}_full_t1 -f python bench.py $SHORT > $out/${tag}_ncu_full_t1.log 2>&1
# a large batch (65536 envs: ray queue, pairs of env groups per CTA), one steady-state launch of 100 steps
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 6 --launch-count 1 \
    -o $out/${tag}_full_big -f python bench.py --envs 65536 --steps-per-launch 100 --steps 300 --warmup 600 --min-timed-steps 300 --min-warmup 600 --e2e-steps 20 --extras 0 --cpu-baseline 0 > $out/${tag}_ncu_full_big.log 2>&1
# the contact path: the "driving" distribution (71 % of car-steps in wall contact), one steady-state launch of 200 steps
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 15 --launch-count 1 \
    -o $out/${tag}_full_drive -f python bench.py --mode 1 --steps-per-launch 200 --steps 200 --warmup 3000 --min-timed-steps 200 --min-warmup 3000 --e2e-steps 20 --extras 0 --cpu-baseline 0 > $out/${tag}_ncu_full_drive.log 2>&1
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench.json; cat $out/${tag}_bench_ref.json
