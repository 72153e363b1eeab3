#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): N-GPU == 1-GPU checks, then the bench line at N GPUs.  G = number of GPUs.
G=${G:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
if [ "${SKIP_CHECK:-0}" != "1" ]; then
timeout -s KILL 600 $TR --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$G.log 2>&1; echo "dist_check exit $?"; grep -h "^rank 0" gpurun_out/dist_check_$G.log | tail -2; tail -3 gpurun_out/dist_check_$G.log
fi
timeout -s KILL 900 $TR --master-port 29512 bench.py --gpus $G --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS} > gpurun_out/bench_${G}gpu.json 2> gpurun_out/bench_${G}gpu.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_${G}gpu.json; tail -5 gpurun_out/bench_${G}gpu.err
if [ -n "${EXTRA}" ]; then eval "${EXTRA}"; fi
