#!/bin/bash
# A/B of the main-pass tiers on one B200 (run under gpurun from the repo root).
# usage: tools/ab_tiers.sh "<tiers>" "<workloads>"
mkdir -p gpurun_out
summ='import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print(sys.argv[1], sys.argv[2], "cells/s %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "clip ms %.3f" % r["avg_launch_ms"], "sm_mhz", d["clocks"]["sm_mhz"] if d.get("clocks") else None, "faces", d.get("checks", {}).get("n_faces"))'
for w in $2; do
for v in $1; do
  TESS_MAIN_TIER=$v TESS_TRACE=1 timeout 300 python bench.py --workload $w --steps 5 --no-cpu-baseline --no-e2e 2>gpurun_out/abt_${w}_$v.err | tee gpurun_out/abt_${w}_$v.json | python -c "$summ" $w $v
  grep "tess trace" gpurun_out/abt_${w}_$v.err | tail -3
done
done
