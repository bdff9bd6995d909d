#!/bin/bash
# A/B of an environment switch on ONE box: scripts/gpu_ab_env.sh VAR A B [repeats]
mkdir -p gpurun_out
for r in $(seq 1 ${4:-2}); do
  for v in "$2" "$3"; do
    env $1=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -c 400 gpurun_out/ab.err
    python - "$1=$v" <<'PY'
import json, sys
d = json.load(open("gpurun_out/ab.json"))
print(f"[{sys.argv[1]}] {d['ms_per_step']:.2f} ms  {d['value']:.1f} clips/s  e2e {d['e2e']['ms_per_step']:.2f} ms  sm {d['clocks']['sm_mhz']} MHz")
PY
  done
done
