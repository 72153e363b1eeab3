#!/bin/bash
# Runs each scoring-kernel test group in its own process with a hard timeout so one hung kernel cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for k in "fwd_train and 1]" "fwd_train and 2]" "fwd_train and 4]" "fwd_eval" "bwd_q" "bwd_i"; do
  tag=$(echo "$k" | tr -c 'a-z0-9_' '_')
  timeout -s KILL 240 python -m pytest tests/test_gpu_score_kernels.py -q -m gpu -k "$k" -x --no-header -p no:cacheprovider > gpurun_out/bringup_$tag.log 2>&1
  echo "== $k -> exit $?" | tee -a gpurun_out/bringup_summary.txt
  tail -5 gpurun_out/bringup_$tag.log
done
