#!/bin/bash
# First multi-GPU visit of the next round: the opt-in look-ahead of the catalog-sharded step.
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/gpu_lookahead_ab.sh 2'
# 1. trajectory equality with the look-ahead ON (tools/dist_check.py honours TCAR_CATALOG_LOOKAHEAD)
# 2. bench A/B: look-ahead off / on (short: no CPU leg, no per-kernel pass, no loop)
G=${1:-2}
mkdir -p gpurun_out
TCAR_CATALOG_LOOKAHEAD=1 timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_look_$G.log 2>&1
echo "dist_check(lookahead) exit $?"; grep "rank 0" gpurun_out/dist_check_look_$G.log | tail -1 | cut -c1-1300
for look in 0 1; do
TCAR_CATALOG_LOOKAHEAD=$look timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 20 --warmup 5 --no_cpu_baseline --no_kernels --loop_sessions 0 > gpurun_out/bench_look${look}_$G.json 2> gpurun_out/bench_look${look}_$G.err
echo "bench look=$look exit $?"; tail -2 gpurun_out/bench_look${look}_$G.err | cut -c1-200
python - $look $G <<'PY'
import json,sys
try:
    d=json.loads(open(f'gpurun_out/bench_look{sys.argv[1]}_{sys.argv[2]}.json').read().strip().splitlines()[-1])
    print("look", sys.argv[1], "train", round(d["value"]), "ms", round(d["ms_per_step"],4), 'e2e', round(d['e2e']['value']), 't20 ms', round(d['t20']['ms_per_step'],4), 'loss', d['loss_last_step'], 'lookahead', d['config']['lookahead'])
except Exception as e:
    print("no bench line:", e)
PY
done
