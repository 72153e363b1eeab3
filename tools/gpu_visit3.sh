#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --no-header -p no:cacheprovider -k "lookahead or determin" 2>&1 | tail -3
summ() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "train", round(d["value"]), "ms", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"]), "t20 ms", round(d["t20"]["ms_per_step"],4), "eval", round(d["eval"]["value"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for pr in 1 0; do
for c in 6 16 64 256; do
TCAR_AHEAD_PRIORITY=$pr TCAR_ADAM_OVERLAP_CTAS=$c timeout -s KILL 400 python bench.py --no_kernels --no_cpu_baseline --loop_sessions 0 > gpurun_out/bench_p${pr}_c$c.json 2> gpurun_out/bench_p${pr}_c$c.err; summ gpurun_out/bench_p${pr}_c$c.json
done
done
