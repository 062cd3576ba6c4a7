#!/bin/bash
# A/B of library build variants on one B200 (run under gpurun from the repo root).
# usage: tools/ab_run.sh "<variants>" "<workloads>" [tier]   (variants: ab_build/libtess_<v>.so, built by tools/ab_build.sh)
mkdir -p gpurun_out
summ='import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print(sys.argv[1], sys.argv[2], "cells/s %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "clip ms %.3f" % r["avg_launch_ms"], "sm_mhz", d["clocks"]["sm_mhz"] if d.get("clocks") else None)'
for w in $2; do
for v in $1; do
  TESS_MAIN_TIER=${3:-thread} TESS_LIB_PATH=$PWD/ab_build/libtess_$v.so timeout 120 python bench.py --workload $w --steps 5 --no-cpu-baseline --no-e2e 2>gpurun_out/ab_${w}_$v.err | tee gpurun_out/ab_${w}_$v.json | python -c "$summ" $w $v
done
done
