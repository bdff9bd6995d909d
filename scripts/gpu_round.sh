#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 600 -x > gpurun_out/pytest_conv.log 2>&1; echo "pytest conv exit $?"; tail -3 gpurun_out/pytest_conv.log
timeout 600 python scripts/bench_conv.py 32 > gpurun_out/bench_conv.log 2>&1; echo "bench_conv exit $?"; tail -32 gpurun_out/bench_conv.log
