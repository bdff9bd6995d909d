#!/bin/bash
# round 2: GPU test suite + graphed / eager bench (run under gpurun)
mkdir -p gpurun_out
rm -f gpurun_out/parity.jsonl
python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err
python bench.py --steps 5 --warmup 3 --graph 0 --no-cpu-baseline --no-extra > gpurun_out/bench_eager.json 2> gpurun_out/bench_eager.err
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench_graph.err; head -c 600 gpurun_out/bench_graph.json; echo; tail -3 gpurun_out/bench_eager.err; head -c 300 gpurun_out/bench_eager.json
