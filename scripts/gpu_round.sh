#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_networks.py -q -m gpu --timeout 600 -x -k "full_size or checkpoint" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_new.log
