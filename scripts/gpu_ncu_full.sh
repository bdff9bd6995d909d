#!/bin/bash
# ncu --set full captures of single launches inside ONE training step (scripts/profile_step.py brackets it with cudaProfilerStart/Stop):
#   prof_tapgemm_fwd48 / prof_tapgemm_dgrad48 : a 48-channel 5x5 SID layer, forward and data gradient
#   prof_tapgemm_fwd96                        : a 96-channel 5x5 ContextAggNet layer, forward
#   prof_wgrad96                              : its weight gradient
mkdir -p gpurun_out
cap() {  # name kernel-regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 --launch-skip $3 --launch-count 1 \
    -f -o gpurun_out/prof_$1 python scripts/profile_step.py 32 > gpurun_out/prof_$1.log 2>&1
  echo "$1 exit $?"
}
cap tapgemm_fwd48 tapgemm_f16 5
cap tapgemm_dgrad48 tapgemm_f16 15
cap tapgemm_fwd96 tapgemm_f16 52
cap wgrad96 wgrad_f16 30
ls -la gpurun_out/*.ncu-rep
