#!/bin/bash
# 8-GPU evidence (run under `gpurun --gpus 8` from the repo root): the north-star target runs.
R=${ROUND:-r02}
G=${GPUS:-8}
mkdir -p gpurun_out
run() {  # workload steps extra-args...
  w=$1; st=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $G --steps $st --warmup 3 --workload $w "$@" \
    2>gpurun_out/${R}_bench_${G}gpu_$w.err | tee gpurun_out/${R}_bench_${G}gpu_$w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$w', 'cells/s %.4g' % d['value'], 'ms/step %.2f' % d['ms_per_step'], 'e2e %.4g' % (d['e2e']['value'] if d.get('e2e') else 0), 'clip ms %.2f' % d['roofline']['avg_launch_ms'], d['checks'])"
}
run uniform10m 5
TESS_SHARD_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $G --steps 3 --warmup 3 --no-e2e 2>&1 | grep "shard trace" | tail -2
run uniform100m 3
run bcc100m 3
run clustered10m 3
