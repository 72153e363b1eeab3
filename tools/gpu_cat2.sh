#!/bin/bash
# multi-GPU visit for the catalog-sharded train step: equality check, then bench in both layouts
G=${1:-2}
STEPS=${2:-20}
mkdir -p gpurun_out
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$G.log 2>&1
echo "dist_check exit $?"; grep "rank " gpurun_out/dist_check_$G.log | tail -8 | cut -c1-1500; tail -12 gpurun_out/dist_check_$G.log | cut -c1-300
for mode in catalog dp; do
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps $STEPS --warmup 5 --train_parallel $mode --no_cpu_baseline --no_kernels --loop_sessions 16384 > gpurun_out/bench_${mode}_$G.json 2> gpurun_out/bench_${mode}_$G.err
echo "bench $mode exit $?"; tail -4 gpurun_out/bench_${mode}_$G.err | cut -c1-300
python - $mode $G <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}.json').read().strip().splitlines()[-1])
    print(sys.argv[1], "train", round(d["value"]), "ms", round(d["ms_per_step"],4), 'e2e', round(d['e2e']['value']), 'eval', round(d['eval']['value']), 't20 ms', round(d['t20']['ms_per_step'],4), 'loop', d['train_loop'] and round(d['train_loop']['value']), 'launches/step', d['launches_per_step'], d['config']['parallelism'][:12], 'loss', d['loss_last_step'])
except Exception as e:
    print("no bench line:", e)
PY
done
