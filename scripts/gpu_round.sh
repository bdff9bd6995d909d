#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_transforms.py -q -m gpu --timeout 600 -x -s -k "longform or denoise" > gpurun_out/pytest_new.log 2>&1; echo "pytest exit $?"; tail -6 gpurun_out/pytest_new.log
for b in 1 8 32 128; do timeout 600 python bench.py --workload infer --batch $b --steps 5 2>gpurun_out/infer_$b.err | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('infer batch', d['config']['batch'], round(d['value'],1), 'clips/s', round(d['ms_per_step'],2), 'ms; e2e', round(d['e2e']['value'],1))"; done
timeout 600 python bench.py --workload infer --batch 16 --length 160000 --steps 3 2>gpurun_out/infer_long.err | cut -c1-400
