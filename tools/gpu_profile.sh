#!/bin/bash
# Run on the GPU box (via gpurun): tests, a bench line, the ncu launch list of the same command and one full capture
# of the rollout kernel.  Outputs land in gpurun_out/<tag>_*.
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $out/${tag}_gpu.txt 2>&1
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 > $out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ncg_step_kernel --launch-skip 4 --launch-count 1 \
    -o $out/${tag}_full -f python bench.py --steps 300 --warmup 300 --e2e-steps 20 --cpu-steps 200 > $out/${tag}_ncu_full.log 2>&1
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_bench.json
