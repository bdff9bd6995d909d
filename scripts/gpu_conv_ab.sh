#!/bin/bash
# conv kernel A/B: correctness of every conv geometry, then per-layer timings; extra arguments are "ENV=value" settings to A/B
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -q -m gpu --timeout 300 -x > gpurun_out/test_gpu_conv.log 2>&1; echo "test_gpu_conv exit $?"; tail -3 gpurun_out/test_gpu_conv.log
timeout 300 python scripts/bench_conv.py 32 all > gpurun_out/bench_conv.log 2>&1; echo "bench_conv exit $?"; cat gpurun_out/bench_conv.log
for kv in "$@"; do
  echo "== $kv"
  env $kv timeout 300 python scripts/bench_conv.py 32 all 2>&1 | tee gpurun_out/bench_conv_$kv.log
done
