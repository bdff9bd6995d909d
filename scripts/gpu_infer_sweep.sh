#!/bin/bash
# BASELINE configs[0]/[3]/[4]: forward-only inference sweep (2 s clips, batch 1..512; 10 s clips x 16), eager vs CUDA graph for the small batches.
mkdir -p gpurun_out
out=gpurun_out/infer_sweep.txt
echo "# python bench.py --workload infer --batch B [--graph 0|1] [--length 160000]   (B200, forward only: STFT->SID->gate->STFT->JointModel->cRM+iSTFT)" > $out
run() {
  timeout 300 python bench.py --workload infer --steps ${STEPS:-10} "$@" 2> gpurun_out/infer.err | python -c "
import json, sys
d = json.loads(sys.stdin.read())
c = d['config']
print(f\"batch {c['batch']:4d}  L {c['samples_per_clip']:6d}  graph {int(c['cuda_graph'])}  {d['ms_per_step']:8.2f} ms/step  {d['value']:8.1f} clips/s   e2e {d['e2e']['value']:8.1f} clips/s\")" >> $out || tail -c 300 gpurun_out/infer.err
}
for b in 1 2 4 8; do run --batch $b --graph 0; run --batch $b --graph 1; done
for b in 16 32 64 128; do run --batch $b --graph 1; done
STEPS=4 run --batch 256 --graph 1
STEPS=3 run --batch 512 --graph 0
STEPS=5 run --batch 16 --length 160000 --graph 0
cat $out
