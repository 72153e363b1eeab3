#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_gemm.py -q -m gpu --no-header -p no:cacheprovider > gpurun_out/gemm_tests.log 2>&1
echo "gemm tests exit $?"; tail -40 gpurun_out/gemm_tests.log | cut -c1-220
