#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_score_kernels test_gpu_gemm test_gpu_parity; do
  timeout -s KILL 600 python -m pytest tests/$f.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/dbg_$f.log 2>&1
  echo "== $f exit $?"; grep -E "passed|failed" gpurun_out/dbg_$f.log | tail -1; grep -E "^FAILED|^ERROR" gpurun_out/dbg_$f.log | head -12
done
