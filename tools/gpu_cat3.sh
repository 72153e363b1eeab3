#!/bin/bash
# catalog-sharded step: equality check + phase probe + short bench at G GPUs
G=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$G.log 2>&1
echo "dist_check exit $?"; grep "rank 0" gpurun_out/dist_check_$G.log | tail -2 | cut -c1-1200; grep -i "error\|Traceback" -A5 gpurun_out/dist_check_$G.log | head -20
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 tools/catalog_probe.py > gpurun_out/probe_$G.log 2>&1
echo "probe exit $?"; grep "^probe" gpurun_out/probe_$G.log | head -2; grep -i "error\|Traceback" -A8 gpurun_out/probe_$G.log | head -20
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 20 --warmup 5 --train_parallel catalog --no_cpu_baseline --no_kernels --loop_sessions 16384 > gpurun_out/bench_catalog_$G.json 2> gpurun_out/bench_catalog_$G.err
echo "bench exit $?"; tail -3 gpurun_out/bench_catalog_$G.err | cut -c1-300
python - $G <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/bench_catalog_{sys.argv[1]}.json').read().strip().splitlines()[-1])
    print("train", round(d["value"]), "ms", round(d["ms_per_step"],4), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']), 't20 ms', round(d['t20']['ms_per_step'],4), 'loop', d['train_loop'] and round(d['train_loop']['value']), d['train_loop'] and d['train_loop'].get('device_sampler',{}).get('value'), 'launches/step', d['launches_per_step'], 'loss', d['loss_last_step'])
except Exception as e:
    print("no bench line:", e)
PY
