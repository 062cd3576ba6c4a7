#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root): full GPU test tier, smoke, bench lines,
# ncu launch list, one full ncu capture of the clip kernel and of the binning kernels.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) | tee gpurun_out/final_tests.txt
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ) | tee -a gpurun_out/final_tests.txt
timeout 600 python bench.py 2>gpurun_out/bench_1gpu.err | tee gpurun_out/bench_1gpu.json | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>gpurun_out/bench_ref.err | tee gpurun_out/bench_ref.json | cut -c1-300
timeout 300 python bench.py --workload clustered10m --steps 3 --no-cpu-baseline 2>gpurun_out/bench_clustered.err | tee gpurun_out/bench_clustered10m.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_launches_10m.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/under_ncu.log 2>&1
# every step launches clip_kernel twice (small configuration, then the medium-configuration redo): 3 warm-up steps = 6 launches to skip
timeout 600 ncu --set full --clock-control none --import-source on -k regex:clip_kernel --launch-skip 6 -c 1 -f -o gpurun_out/r01_clip_10m python bench.py --steps 1 --no-cpu-baseline --no-e2e > gpurun_out/under_ncu2.log 2>&1
timeout 200 ncu --set full --clock-control none -k regex:"cell_histogram|scan_kernel|scatter_records|rank_fix|bounds_partial" -c 10 -f -o gpurun_out/r01_binning_after python tools/init_only.py 10000000 > gpurun_out/under_ncu3.log 2>&1
ls -la gpurun_out | tail -20
