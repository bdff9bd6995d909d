#!/bin/bash
# ncu --set full of single launches of ONE layer of scripts/bench_conv.py (in-step call forms): scripts/gpu_ncu_layer.sh <layer> [roles...]
# roles: fwd dgrad wgrad rowfwd rowdgrad (default: fwd).  Reports land in gpurun_out/prof_<layer>_<role>.ncu-rep
mkdir -p gpurun_out
L=$1; shift
for role in ${@:-fwd}; do
  case $role in
    fwd) K=tapgemm_f16; SKIP=5;;
    dgrad) K=tapgemm_f16; SKIP=12;;
    wgrad) K=wgrad_f16; SKIP=4;;
    rowfwd) K=rowconv_f16; SKIP=4;;
    rowdgrad) K=rowconv_f16; SKIP=10;;
  esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K --launch-skip $SKIP --launch-count 1 \
    -f -o gpurun_out/prof_${L}_$role python scripts/bench_conv.py 32 $L > gpurun_out/prof_${L}_$role.log 2>&1
  echo "$L $role exit $?"
done
