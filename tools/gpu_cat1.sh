#!/bin/bash
# 1-GPU visit for the catalog-sharded train step: new parity tests first (fail fast), then the whole GPU suite, smoke
# and a short bench.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --no-header -p no:cacheprovider -k "catalog or range or accumulate" > gpurun_out/pytest_catalog.log 2>&1
echo "catalog pytest exit $?"; tail -40 gpurun_out/pytest_catalog.log | cut -c1-400
timeout -s KILL 900 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --deselect tests/test_gpu_parity.py::test_catalog_sharded_train_step_equals_plain_step > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no_cpu_baseline --no_kernels --loop_sessions 0 > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench exit $?"; tail -2 gpurun_out/bench_short.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_short.json').read().strip().splitlines()[-1])
print("train", d["value"], "ms", d["ms_per_step"], 'e2e', d['e2e']['value'], 'eval', d['eval']['value'], 't20', d['t20']['ms_per_step'])
PY
