#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_transforms.py -q -m gpu --timeout 300 -x -k "stft or longform or gat" > gpurun_out/pytest_stft.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_stft.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stft_tc python scripts/stft_probe.py 32 2>&1 | grep -E "gpu__time_duration" | head -3
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:stft_tc python scripts/stft_probe.py 512 2>&1 | grep -E "gpu__time_duration" | head -3
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench exit $?"; tail -c 300 gpurun_out/bench_b32.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print(d['value'], d['ms_per_step']); print(d['kernels']['stft'])"
