#!/bin/bash
# A/B of the two-physics-warp CTA shape (NCG_PHYS_WARPS=1/2) + the whole GPU suite with it forced on
tag=${1:-pw}
out=gpurun_out; mkdir -p $out
NCG_PHYS_WARPS=2 timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_pw2.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest_pw2.log
tail -8 $out/${tag}_pytest_pw2.log
run() {
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
for pw in ${PWLIST:-1 2}; do
  run e8192_pw$pw NCG_PHYS_WARPS=$pw -- --envs 8192 --steps 2000 --warmup 500 --steps-per-launch 500
  run e16384_pw$pw NCG_PHYS_WARPS=$pw -- --envs 16384 --steps 1000 --warmup 300 --steps-per-launch 250
  run e65536_pw$pw NCG_PHYS_WARPS=$pw -- --envs 65536 --steps 400 --warmup 100 --steps-per-launch 100
  run e65536all_pw$pw NCG_PHYS_WARPS=$pw -- --envs 65536 --track all --steps 400 --warmup 100 --steps-per-launch 100
  run c10_pw$pw NCG_PHYS_WARPS=$pw -- --envs 8192 --cars 10 --track talladega --steps 300 --warmup 100 --steps-per-launch 100
done
