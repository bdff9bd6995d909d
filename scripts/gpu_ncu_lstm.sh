#!/bin/bash
# ncu --set full captures of the cluster LSTM kernels inside ONE training step (the JointModel's T = 203, H = 200 launches).
mkdir -p gpurun_out
cap() {  # name kernel-regex skip
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 --launch-skip $3 --launch-count 1 \
    -f -o gpurun_out/prof_$1 python scripts/profile_step.py 32 > gpurun_out/prof_$1.log 2>&1
  echo "$1 exit $?"
}
cap lstm_fwd_cluster lstm_fwd_cluster 1
cap lstm_bwd_cluster lstm_bwd_cluster 0
ls -la gpurun_out/prof_lstm*.ncu-rep
