#!/bin/bash
# `ncu --set full` capture of every hot kernel of ONE train step (T = 20) + ONE eval step (bench.py --profile_region)
mkdir -p gpurun_out
K='regex:score_fwd_pair|score_bwd_i|adam_item|gather_fwd|pool_fwd|pool_bwd|scatter_|table_grads|gemm_tf32_kernel|eval_topk|build_query|col_jobs|update_norms|adam_small|neg_loss|loss_combine|score_bwd_finish|ce_finish|prep_weights'
timeout -s KILL 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "$K" -f -o /tmp/prof_step \
  python bench.py --profile_region > gpurun_out/ncu_full.log 2>&1
echo "ncu A exit $?"; tail -2 gpurun_out/ncu_full.log
# score_bwd_q fails to launch under the full set's instrumentation (it owns all 227 KB of shared memory): lighter sections
timeout -s KILL 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy --clock-control none --profile-from-start off \
  -k regex:score_bwd_q_kernel -f -o /tmp/prof_bwd_q python bench.py --profile_region > gpurun_out/ncu_bwd_q.log 2>&1
echo "ncu B exit $?"; tail -2 gpurun_out/ncu_bwd_q.log
# gpurun copies back at most 64 MiB and drops EVERYTHING when gpurun_out/ is larger: the reports stay in /tmp (a killed
# run then leaves nothing oversized behind), only the exported raw pages travel
for r in prof_step prof_bwd_q; do
  ncu -i /tmp/$r.ncu-rep --page raw --csv > gpurun_out/${r}_raw.csv 2>/dev/null
done
ls -la gpurun_out/prof_* /tmp/prof_*.ncu-rep
