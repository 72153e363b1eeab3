#!/bin/bash
# visit: parity tests + bench with / without the one-batch look-ahead, sweep of the Adam grid used under overlap
mkdir -p gpurun_out
bash tools/gpu_quick.sh
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "train", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "t20 ms", round(d["t20"]["ms_per_step"],4), "eval", round(d["eval"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
timeout -s KILL 400 python bench.py --no_lookahead --no_kernels --no_cpu_baseline --loop_sessions 0 > gpurun_out/bench_nolook.json 2> gpurun_out/bench_nolook.err; summ gpurun_out/bench_nolook.json
for c in 1 3 4 6; do
TCAR_ADAM_OVERLAP_CTAS=$c timeout -s KILL 400 python bench.py --no_kernels --no_cpu_baseline --loop_sessions 0 > gpurun_out/bench_ctas$c.json 2> gpurun_out/bench_ctas$c.err; summ gpurun_out/bench_ctas$c.json
done
