#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_score_kernels.py -q -m gpu -k "fwd" -x --no-header -p no:cacheprovider > gpurun_out/fwd_tests.log 2>&1
echo "fwd tests exit $?"; tail -15 gpurun_out/fwd_tests.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/bench_fwd.json 2> gpurun_out/bench_fwd.err; echo "bench exit $?"; tail -3 gpurun_out/bench_fwd.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_fwd.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['eval']['value'], d['clocks'])
for k,v in d['kernels'].items(): print(f"{k:20s} {v['ms']*1000:8.1f} us  {v['achieved']:8.1f} {v['unit']} frac {v['frac']:.3f}")
PY
