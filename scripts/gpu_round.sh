#!/bin/bash
# Profiling visit: smoke, per-layer table, launch list of one step, --set full captures of the tensor-core kernels, default bench.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 600 python scripts/bench_conv.py 32 > gpurun_out/bench_conv.log 2>&1; echo "bench_conv exit $?"; tail -2 gpurun_out/bench_conv.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py 32 > gpurun_out/launches.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:tapgemm -s 60 -c 2 -o gpurun_out/prof_tapgemm -f python scripts/profile_step.py 32 > gpurun_out/prof_tapgemm.log 2>&1; echo "ncu tapgemm exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:wgrad_tf32 -s 30 -c 2 -o gpurun_out/prof_wgrad -f python scripts/profile_step.py 32 > gpurun_out/prof_wgrad.log 2>&1; echo "ncu wgrad exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stft_tc -s 1 -c 1 -o gpurun_out/prof_stft -f python scripts/stft_probe.py 512 > gpurun_out/prof_stft.log 2>&1; echo "ncu stft exit $?"
timeout 900 python bench.py > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench exit $?"; tail -c 300 gpurun_out/bench_b32.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"; cut -c1-300 gpurun_out/bench_ref.json
