#!/bin/bash
# parity tests + bench without the CPU baseline; prints a compact summary
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -6 gpurun_out/pytest_gpu.log
timeout -s KILL 600 python bench.py --steps 20 --warmup 5 --no_cpu_baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; tail -3 gpurun_out/bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print("t20", d["t20"]); print("loop", d["train_loop"]); print("train", d["value"], "ms", d["ms_per_step"], 'e2e', d['e2e']['value'], 'eval', d['eval']['value'], d['clocks'], 'launches/step', d['launches_per_step'])
for k,v in d['kernels'].items(): print(f"{k:20s} {v['ms']*1000:8.1f} us  {v['achieved']:8.1f} {v['unit']} frac {v['frac']:.3f}")
PY
if [ "$1" == "ncu" ]; then
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_kernels > gpurun_out/bench_ncu.log 2>&1; echo "ncu exit $?"
fi
