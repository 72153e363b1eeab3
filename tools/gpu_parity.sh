#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/parity.log 2>&1
echo "parity exit $?"; tail -40 gpurun_out/parity.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -5 gpurun_out/smoke.log
