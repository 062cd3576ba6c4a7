#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root): full GPU test tier, smoke, bench lines,
# ncu launch list, one full ncu capture of the shipped clip kernel, compute-sanitizer runs.
R=${ROUND:-r02}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/${R}_gpu_tests.txt
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee -a gpurun_out/${R}_gpu_tests.txt
timeout 600 python bench.py 2>gpurun_out/bench_1gpu.err | tee gpurun_out/${R}_bench_1gpu.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/${R}_bench_ref.json | cut -c1-300
timeout 300 python bench.py --workload clustered10m --steps 3 --no-cpu-baseline 2>gpurun_out/bench_clustered.err | tee gpurun_out/${R}_bench_1gpu_clustered10m.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_10m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/under_ncu.log 2>&1
# every step launches the main clip kernel once (clip_kernel<SmallCfg,0,0>), then the redo tiers: skip the 3 warm-up steps
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"clip_kernel.*SmallCfg, .bool.0, .bool.0" --launch-skip 3 -c 1 -f -o gpurun_out/${R}_clip_10m python bench.py --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/under_ncu2.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py all 2>&1 | tail -6; echo "memcheck exit $?" ) | tee gpurun_out/${R}_sanitizer.txt
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py warp 3000 > gpurun_out/${R}_racecheck_full.txt 2>&1; grep -E "Error:|Warning:|SUMMARY|sanitize_run ok" gpurun_out/${R}_racecheck_full.txt | sed "s/+0x[0-9a-f]*//" | sort | uniq -c | sort -rn | head -30; echo "racecheck exit $?" ) | tee -a gpurun_out/${R}_sanitizer.txt
ls -la gpurun_out | tail -12
