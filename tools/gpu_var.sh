#!/bin/bash
# bench variants given as "NAME:ENV=VAL,ENV=VAL:extra bench args" ...
tag=$1; shift
out=gpurun_out; mkdir -p $out
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; envs=${rest%%:*}; args=${rest#*:}; [ "$args" = "$rest" ] && args=""
  env $(echo $envs | tr ',' ' ') timeout 300 python bench.py --steps 2000 --warmup 600 --e2e-steps 200 --cpu-steps 200 --sweep 0 $args > $out/${tag}_$name.json 2> $out/${tag}_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_$name.json")); print("$name: value %.1fM e2e %.1fM us/step %.2f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"]*1e3), {k:v for k,v in d["counters"].items() if v})
except Exception as e: print("$name failed", e); print(open("$out/${tag}_$name.err").read()[-1500:])
PY
done
