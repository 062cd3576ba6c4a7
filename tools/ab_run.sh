#!/bin/bash
# A/B of clip-kernel build variants on one B200 (run under gpurun from the repo root).
# usage: tools/ab_run.sh "<variants for uniform1m>" "<variants for clustered10m>"
mkdir -p gpurun_out
summ='import json,sys
d=json.loads(sys.stdin.read())
r=d["roofline"]
print(sys.argv[1], sys.argv[2], "cells/s %.4g" % d["value"], "ms/step %.3f" % d["ms_per_step"], "clip ms %.3f" % r["avg_launch_ms"], "sm_mhz", d["clocks"]["sm_mhz"] if d.get("clocks") else None)'
for v in $1; do
  TESS_LIB_PATH=$PWD/ab_build/libtess_$v.so timeout 240 python bench.py --workload uniform1m --steps 10 --no-cpu-baseline --no-e2e 2>gpurun_out/ab_u1m_$v.err | tee gpurun_out/ab_u1m_$v.json | python -c "$summ" uniform1m $v
done
for v in $2; do
  TESS_LIB_PATH=$PWD/ab_build/libtess_$v.so TESS_TRACE=1 timeout 300 python bench.py --workload clustered10m --steps 3 --no-cpu-baseline --no-e2e 2>gpurun_out/ab_c10m_$v.err | tee gpurun_out/ab_c10m_$v.json | python -c "$summ" clustered10m $v
  grep "tess trace" gpurun_out/ab_c10m_$v.err | tail -12
done
