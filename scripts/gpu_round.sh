#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench exit $?"; tail -c 300 gpurun_out/bench_b32.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print(d['value'], d['ms_per_step'], d['gpu_launches'])"
