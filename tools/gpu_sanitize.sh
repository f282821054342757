#!/bin/bash
# compute-sanitizer over every launch shape of the step kernel (run on the GPU box); tails land in gpurun_out/sanitize_*.log
out=gpurun_out
mkdir -p $out
run() {  # tool, tag, env...
  tool=$1; tag=$2; shift 2
  env "$@" compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_small.py > $out/sanitize_${tool}_${tag}.full 2>&1
  rc=$?
  { echo "### compute-sanitizer --tool $tool  [$*]  exit code $rc"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize workload ok|Error|hazard" $out/sanitize_${tool}_${tag}.full | head -20; } >> $out/sanitize_summary.log
  rm -f $out/sanitize_${tool}_${tag}.full
}
: > $out/sanitize_summary.log
for tool in memcheck racecheck synccheck; do
  run $tool default NCG_X=0
  run $tool queue_rpl4 NCG_RAY_QUEUE=1 NCG_RAYS_PER_LANE=4
  run $tool pair NCG_PHYS_WARPS=2 NCG_RAY_QUEUE=1 NCG_RAYS_PER_LANE=4
  run $tool spread NCG_PHYS_WARPS=4
  run $tool nostage NCG_NO_STAGE=1 NCG_MIN_BLOCKS=3 NCG_RAYS_PER_LANE=4
done
cat $out/sanitize_summary.log
