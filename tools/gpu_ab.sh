#!/bin/bash
# A/B of build-time variants on ONE box: tools/gpu_ab.sh "<EXTRA flags A>" "<EXTRA flags B>" ...
mkdir -p gpurun_out
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "train ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "t20 ms", round(d["t20"]["ms_per_step"],4), "eval", round(d["eval"]["value"]), d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
i=0
for rep in 1 2; do
for ex in "$@"; do
  i=$((i+1))
  make -C session-based-news-recommendation_b200/csrc -B -j8 EXTRA="$ex" > gpurun_out/build_$i.log 2>&1 || { echo "build failed: $ex"; tail -5 gpurun_out/build_$i.log; continue; }
  timeout -s KILL 400 python bench.py --no_kernels --no_cpu_baseline --loop_sessions 0 > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err
  echo "[$ex]"; summ gpurun_out/ab_$i.json
done
done
