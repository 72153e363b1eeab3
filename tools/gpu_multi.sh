#!/bin/bash
# multi-GPU visit: distributed equality check + bench at N GPUs
G=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/dist_check_$G.log 2>&1
echo "dist_check exit $?"; grep "rank " gpurun_out/dist_check_$G.log | tail -8; tail -3 gpurun_out/dist_check_$G.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/bench_$G.json 2> gpurun_out/bench_$G.err
echo "bench exit $?"; tail -c 1500 gpurun_out/bench_$G.json; tail -5 gpurun_out/bench_$G.err
