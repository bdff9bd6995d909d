#!/bin/bash
# round 2 profiles: (1) launch list of ONE eager training step, (2) ncu --set full of single launches inside the step, (3) per-layer conv table
mkdir -p gpurun_out
rm -f gpurun_out/prof_*.ncu-rep
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python scripts/profile_step.py 32 > gpurun_out/launches.log 2>&1
echo "launch list exit $?"
cap() {  # name kernel-regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 --launch-skip $3 --launch-count 1 \
    -f -o gpurun_out/prof_$1 python scripts/profile_step.py 32 > gpurun_out/prof_$1.log 2>&1
  echo "$1 exit $?"
}
cap tapgemm_fwd48 tapgemm_f16 5
cap tapgemm_fwd96 tapgemm_f16 52
cap wgrad96 wgrad_f16 30
cap bn_bwd_reduce bn_bwd_reduce_h 30
cap bn_bwd_apply bn_bwd_apply_h 30
cap bn_act bn_act_h_kernel 30
timeout 300 python scripts/bench_conv.py > gpurun_out/bench_conv.log 2>&1; echo "bench_conv exit $?"
ls -la gpurun_out/*.ncu-rep
