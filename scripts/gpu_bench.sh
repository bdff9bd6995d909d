#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_networks.py -q -m gpu --timeout 300 -s > gpurun_out/test_gpu_networks.log 2>&1; echo "networks exit $?"
timeout 600 python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b8.json 2> gpurun_out/bench_b8.err; echo "bench b8 exit $?"; tail -c 600 gpurun_out/bench_b8.err
timeout 900 python bench.py --batch 32 --steps 3 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench b32 exit $?"; tail -c 600 gpurun_out/bench_b32.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
