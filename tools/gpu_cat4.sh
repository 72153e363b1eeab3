#!/bin/bash
# shortest multi-GPU confirmation: equality check + short catalog bench
G=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$G.log 2>&1
echo "dist_check exit $?"; grep "rank 0" gpurun_out/dist_check_$G.log | tail -1 | cut -c1-1200
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 20 --warmup 5 --no_cpu_baseline --no_kernels --loop_sessions 0 > gpurun_out/bench_catalog_$G.json 2> gpurun_out/bench_catalog_$G.err
echo "bench exit $?"; tail -2 gpurun_out/bench_catalog_$G.err | cut -c1-200
python - $G <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/bench_catalog_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print("train", round(d["value"]), "ms", round(d["ms_per_step"],4), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']), 't20 ms', round(d['t20']['ms_per_step'],4), 'launches/step', d['launches_per_step'], 'loss', d['loss_last_step'])
PY
