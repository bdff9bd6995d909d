#!/bin/bash
# launch list of ONE eager training step (ncu gpu__time_duration per launch) -> gpurun_out/launches.csv
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python scripts/profile_step.py 32 > gpurun_out/launches.log 2>&1
echo "launch list exit $?"
