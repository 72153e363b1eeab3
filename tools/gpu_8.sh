#!/bin/bash
# 8-GPU visit: correctness at 8 ranks, the default bench line, probes, and the data-parallel line.
G=${G:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
timeout -s KILL 300 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$G.log 2>&1; echo "dist_check exit $?"; grep -c '"ok": true}$' gpurun_out/dist_check_$G.log
fi
timeout -s KILL 500 $TR --master-port 29512 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err; echo "bench exit $?"; python tools/show_bench.py gpurun_out/bench_${G}gpu.json; tail -3 gpurun_out/bench_${G}gpu.err
timeout -s KILL 200 $TR --master-port 29514 tools/catalog_probe.py 2>&1 | grep "^probe" | head -1 > gpurun_out/catalog_probe_$G.log; cat gpurun_out/catalog_probe_$G.log
if [ "${SKIP_EVAL_PROBE:-0}" != "1" ]; then
timeout -s KILL 200 $TR --master-port 29513 tools/eval_round_probe.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" > gpurun_out/eval_probe_$G.log; cat gpurun_out/eval_probe_$G.log
fi
if [ "${SKIP_DP:-0}" != "1" ]; then
timeout -s KILL 300 $TR --master-port 29515 bench.py --gpus $G --steps 20 --warmup 5 --train_parallel dp --no_kernels --no_cpu_baseline --loop_sessions 0 > gpurun_out/bench_${G}gpu_dp.json 2> gpurun_out/bench_${G}gpu_dp.err; echo "bench dp exit $?"; python tools/show_bench.py gpurun_out/bench_${G}gpu_dp.json; tail -3 gpurun_out/bench_${G}gpu_dp.err
fi
if [ -n "${EXTRA}" ]; then eval "${EXTRA}"; fi
