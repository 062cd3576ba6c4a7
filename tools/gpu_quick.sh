#!/bin/bash
# quick GPU check (run from the repo root under gpurun): parity tests without the 10M case + a 1M bench summary
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_parity.py::test_config3_ten_million_uniform 2>&1 | tail -4
timeout 300 python bench.py --workload uniform1m --steps 3 --no-cpu-baseline 2>gpurun_out/b1.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('cells/s', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'clip ms', d['roofline']['avg_launch_ms'], 'roofline', d['roofline']['bound'], d['roofline']['frac'])
"
tail -2 gpurun_out/b1.err
