#!/bin/bash
# conv kernel A/B: correctness of every conv geometry, then per-layer timings (default plan and forced K chunks)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 -x > gpurun_out/test_gpu_conv.log 2>&1; echo "test_gpu_conv exit $?"; tail -3 gpurun_out/test_gpu_conv.log
timeout 300 python scripts/bench_conv.py 32 all > gpurun_out/bench_conv.log 2>&1; echo "bench_conv exit $?"; cat gpurun_out/bench_conv.log
for cbe in "$@"; do
  echo "== SOS_FORCE_CBE=$cbe"
  SOS_FORCE_CBE=$cbe timeout 300 python scripts/bench_conv.py 32 n48 2>&1 | tee gpurun_out/bench_conv_cbe$cbe.log
done
