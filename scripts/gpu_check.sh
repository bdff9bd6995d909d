#!/bin/bash
# Runs the GPU test files one process each (a device trap in one must not poison the others); logs under gpurun_out/.
mkdir -p gpurun_out
for s in "$@"; do
  n=$(basename $s .py)
  timeout 600 python $s > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary.txt
done
for f in tests/test_gpu_transforms.py tests/test_gpu_conv.py tests/test_gpu_networks.py; do
  n=$(basename $f .py)
  timeout 900 python -m pytest $f -q -m gpu --timeout 300 -s > gpurun_out/$n.log 2>&1
  echo "$n exit $?" | tee -a gpurun_out/summary.txt
  grep -E "passed|failed" gpurun_out/$n.log | tail -3
done
