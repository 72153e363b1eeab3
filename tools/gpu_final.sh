#!/bin/bash
# last visit of a round: the new kernel tests first, then the whole GPU suite and smoke
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_score_kernels.py -q -m gpu --no-header -p no:cacheprovider -k "groups" > gpurun_out/pytest_groups.log 2>&1
echo "groups pytest exit $?"; tail -30 gpurun_out/pytest_groups.log | cut -c1-300
timeout -s KILL 400 python -m pytest tests -q -m gpu --no-header -p no:cacheprovider --deselect tests/test_gpu_score_kernels.py::test_score_groups_match_single_group_calls > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-300
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
