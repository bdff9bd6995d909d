#!/bin/bash
# ncu --set full capture of the STFT kernel at 128 signals (one launch)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:stft_f16 --launch-skip 3 --launch-count 1 -f -o gpurun_out/prof_stft_f16 \
  python scripts/bench_transforms.py 128 > gpurun_out/prof_stft_f16.log 2>&1
echo "exit $?"
ncu -i gpurun_out/prof_stft_f16.ncu-rep --page raw --csv > gpurun_out/prof_stft_f16_raw.csv
python - <<'PY'
import csv
rows=list(csv.reader(open("gpurun_out/prof_stft_f16_raw.csv")))
hdr,units,vals=rows[0],rows[1],rows[2]
want=["gpu__time_duration.sum","sm__cycles_elapsed.avg","sm__pipe_tensor","tensor","sm__throughput","lts__throughput","dram__throughput","l1tex__data_pipe","l1tex__throughput","smsp__inst_executed.sum","sm__inst_executed","smsp__issue_active","warp_issue_stalled","smsp__average_warp","dram__bytes","l1tex__m_xbar2l1tex","lsu","shared","smsp__cycles_active"]
for h,u,v in zip(hdr,units,vals):
    if any(w in h for w in want) and ("pct" in h or "sum" in h or "avg" in h or "ratio" in h):
        print(f"{h:110s} {v:>18s} {u}")
PY
