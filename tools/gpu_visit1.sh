#!/bin/bash
# visit: parity tests + bench (no CPU leg) + full ncu capture of one train step and one eval step (report comes back)
bash tools/gpu_quick.sh
K='regex:gather_fwd|pool_fwd|pool_bwd|scatter_|table_partial|table_finish|gemm_tf32_kernel|eval_topk|build_query|score_bwd_finish|adam_item|score_bwd_i_kernel|score_fwd_pair|neg_loss|act_bwd|ce_finish'
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -f -o gpurun_out/prof_step \
  python bench.py --profile_region > gpurun_out/ncu_full.log 2>&1
echo "ncu A exit $?"; tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
