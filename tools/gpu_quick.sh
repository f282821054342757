#!/bin/bash
# quick GPU check: parity tests + short bench lines for a few variants
tag=${1:-q}
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> $out/${tag}_pytest.log
tail -15 $out/${tag}_pytest.log
for rpl in 2 4 1; do
  NCG_RAYS_PER_LANE=$rpl timeout 300 python bench.py --steps 2000 --warmup 600 --e2e-steps 300 --cpu-steps 200 > $out/${tag}_bench_rpl$rpl.json 2> $out/${tag}_bench_rpl$rpl.err
  python - <<PY
import json
try:
    d=json.load(open("$out/${tag}_bench_rpl$rpl.json")); print("rpl=$rpl value %.1fM e2e %.1fM ms/step %.4f"%(d["value"]/1e6,d["e2e"]["value"]/1e6,d["ms_per_step"]), d["counters"])
except Exception as e: print("rpl=$rpl failed", e); print(open("$out/${tag}_bench_rpl$rpl.err").read()[-2000:])
PY
done
