"""Turn the ncu outputs of a gpurun visit (gpurun_out/) into the tracked summaries under profiles/.
  launches.csv (gpu__time_duration.sum per launch of ONE training step)  -> profiles/<tag>_launches.txt (+ .csv.gz)
  prof_*.ncu-rep (--set full captures)                                   -> profiles/<tag>_<name>_ncu.txt, profiles/ncu_traffic.json
usage: python scripts/summarize_profiles.py r01"""
import collections
import csv
import gzip
import json
import os
import re
import shutil
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
os.makedirs("profiles", exist_ok=True)

src = "gpurun_out/launches.csv"
if os.path.exists(src):
    lines = open(src).readlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for row in csv.DictReader(lines[start:]):
        if row["Metric Name"] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "")[:90]
        v = float(row["Metric Value"].replace(",", "")) / 1e6
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    with open(f"profiles/{tag}_launches.txt", "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python scripts/profile_step.py 32\n")
        f.write(f"# ONE training step (batch 32, BASELINE configs[1]); per-launch times are cold-cache and serialised: compare SHARES\n")
        f.write(f"# total {tot:.2f} ms over {sum(n for n, _ in agg.values())} launches\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.3f} ms {100 * t / tot:6.2f}%  n={n:5d}  avg {1e3 * t / n:9.1f} us  {k}\n")
    with open(src, "rb") as fi, gzip.open(f"profiles/{tag}_launches.csv.gz", "wb") as fo:
        shutil.copyfileobj(fi, fo)
    print("wrote", f"profiles/{tag}_launches.txt")

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_uniform.sum", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        # shared-memory pipe: wavefronts read by the tensor core, data-bank reads / writes (TMA fills, staging), bytes TMA loaded
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "l1tex__data_bank_reads.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_writes.avg.pct_of_peak_sustained_elapsed", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum"]
traffic = {}
try:                                   # keep the entries of kernels that were not captured again this time
    traffic = json.load(open("profiles/ncu_traffic.json"))
except (OSError, ValueError):
    pass
fresh = set()
for rep in sorted(f for f in os.listdir("gpurun_out") if f.endswith(".ncu-rep")):
    name = rep[:-len(".ncu-rep")].replace("prof_", "")
    raw = subprocess.run(["ncu", "-i", os.path.join("gpurun_out", rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    with open(f"profiles/{tag}_{name}_ncu.txt", "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on (gpurun_out/{rep}); selected raw metrics per captured launch\n")
        for r in rows[2:]:
            d = {h: (v, u) for h, u, v in zip(hdr, units, r)}
            f.write(f"\n== {d['Kernel Name'][0][:100]}  grid {d.get('launch__grid_size', ('?',))[0]}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"   {k:85s} {d[k][0]:>16s} {d[k][1]}\n")
            kn = re.sub(r"\(.*", "", d["Kernel Name"][0]).replace("<unnamed>::", "")
            def num(k):
                v, u = d[k]
                v = float(v.replace(",", ""))
                return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            if "dram__bytes_read.sum" in d:
                t = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
                if (kn not in fresh or t > traffic[kn]["dram_bytes_per_launch"]) and not (kn in traffic and kn not in fresh and "largest captured" not in traffic[kn].get("note", "")):
                    fresh.add(kn)
                    traffic[kn] = {"dram_bytes_per_launch": t, "duration_ms": float(d["gpu__time_duration.sum"][0].replace(",", "")),
                                   "note": f"largest captured launch of {kn} ({tag}, ncu --set full, profiles/{tag}_{name}_ncu.txt)"}
    print("wrote", f"profiles/{tag}_{name}_ncu.txt")
if traffic:
    json.dump(traffic, open("profiles/ncu_traffic.json", "w"), indent=1)
for f in ("bench_conv.log",):
    if os.path.exists("gpurun_out/" + f):
        shutil.copy("gpurun_out/" + f, f"profiles/{tag}_{f.replace('.log', '.txt')}")
