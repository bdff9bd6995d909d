#!/bin/bash
# round 2: GPU test suite + graphed / eager bench (run under gpurun)
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err

tail -8 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_graph.err; head -c 600 gpurun_out/bench_graph.json; echo; tail -3 gpurun_out/bench_eager.err; head -c 300 gpurun_out/bench_eager.json
