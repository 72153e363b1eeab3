#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (own arm + reference arm), ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout -s KILL 1500 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider ${PYTEST_ARGS:--x} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
if [ "${SKIP_BENCH:-0}" != "1" ]; then
timeout -s KILL 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 1500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
fi
if [ "${WITH_REFERENCE:-0}" = "1" ]; then
timeout -s KILL 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference exit $?"; tail -c 600 gpurun_out/bench_reference.json
fi
if [ "${WITH_NCU:-0}" = "1" ]; then
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_kernels --loop_sessions 0 > gpurun_out/bench_ncu.log 2>&1; echo "ncu exit $?"
fi
