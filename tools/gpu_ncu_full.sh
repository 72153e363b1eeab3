#!/bin/bash
# one `ncu --set full` capture of the hot kernels of the second train step
mkdir -p gpurun_out
timeout -s KILL 1200 ncu --set full --clock-control none --import-source on \
  -k 'regex:gemm_tf32_kernel|score_fwd_pair|score_bwd_q_kernel|score_bwd_i_kernel|adam_item|gather_fwd|pool_fwd|pool_bwd|scatter_accum|small_table' \
  --launch-skip ${1:-34} --launch-count ${2:-34} -f -o gpurun_out/prof_step \
  python bench.py --steps 1 --warmup 2 --no_cpu_baseline --no_kernels > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/prof_step.ncu-rep
