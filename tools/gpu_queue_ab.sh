#!/bin/bash
# A/B of the ray queue (NCG_RAY_QUEUE=0/1) at several batch shapes + parity tests with the queue forced on
tag=${1:-qab}
out=gpurun_out; mkdir -p $out
NCG_RAY_QUEUE=1 timeout 700 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -x -q > $out/${tag}_pytest_q1.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_q1.log
tail -8 $out/${tag}_pytest_q1.log
run() {  # name, env assignments..., -- bench args
  name=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --sweep 0 --e2e-steps 200 --cpu-steps 100 "$@" > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); print("$name value %.1fM ms/step %.4f tests/carstep %.1f"%(d["value"]/1e6,d["ms_per_step"],d["counters"]["ray_tests"]/d["counters"]["car_steps"]))
except Exception as e: print("$name failed", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
}
for q in ${QLIST:-0 1}; do
  run e4096_q$q NCG_RAY_QUEUE=$q -- --steps 4000 --warmup 1000
  run e8192_q$q NCG_RAY_QUEUE=$q -- --envs 8192 --steps 2000 --warmup 500 --steps-per-launch 500
  run e16384_q$q NCG_RAY_QUEUE=$q -- --envs 16384 --steps 1000 --warmup 300 --steps-per-launch 250
  run e65536_q$q NCG_RAY_QUEUE=$q -- --envs 65536 --steps 400 --warmup 100 --steps-per-launch 100
  run e65536all_q$q NCG_RAY_QUEUE=$q -- --envs 65536 --track all --steps 400 --warmup 100 --steps-per-launch 100
  run c10_q$q NCG_RAY_QUEUE=$q -- --envs 8192 --cars 10 --track talladega --steps 300 --warmup 100 --steps-per-launch 100
done
run e4096_q1_rpl4 NCG_RAY_QUEUE=1 NCG_RAYS_PER_LANE=4 -- --steps 4000 --warmup 1000
run e65536_q1_rpl2 NCG_RAY_QUEUE=1 NCG_RAYS_PER_LANE=2 -- --envs 65536 --steps 400 --warmup 100 --steps-per-launch 100
