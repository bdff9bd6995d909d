"""GPU timeline of one training step through torch.profiler (CUPTI): busy time per stream, union busy time, idle gaps, and the
largest gaps with the kernels around them.  `python scripts/timeline.py graph`: the graphed step (agent.GraphedTrainStep)."""
import sys
import torch
from torch.profiler import profile, ProfilerActivity
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import agent as ag, ops, tools, transform
import bench

B = 32
ops.init()
dev = torch.device("cuda:0")
torch.manual_seed(0)
sid = ag.SIDAgent(ag.default_config(model="sid"))
torch.manual_seed(1)
joint = ag.MyAgent(ag.default_config(model="joint", sr=bench.SR, fps=bench.FPS))
d = {k: torch.from_numpy(v).to(dev) for k, v in bench.synth_batch(B).items()}
ratio = bench.SR / bench.FPS


def step():
    gated = tools.gate_noise(d["mixed"], ratio, d["bits"])
    spec = transform.stft_batch(torch.cat([d["mixed"], gated, d["clean"], d["full_noise"]]))
    mixed, noise, clean, full = spec[:B], spec[B:2 * B], spec[2 * B:3 * B], spec[3 * B:]
    sid.train_func({"audio": mixed, "label": d["label"]})
    joint.train_func({"mixed": mixed, "noise": noise, "clean": clean, "full_noise": full})
    return transform.istft_batch(joint.last_rec.detach())


if len(sys.argv) > 1 and sys.argv[1] == "graph":          # the step as bench.py times it: two CUDA-graph replays
    gstep = ag.GraphedTrainStep(sid, joint, B, d["mixed"].shape[1], bench.SR, bench.FPS)
    step = lambda: gstep(d["mixed"], d["clean"], d["full_noise"], d["bits"], d["label"])
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(2):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = sorted([(e.time_range.start, e.time_range.end, e.name) for e in ev if e.time_range.end > e.time_range.start])
t0, t1 = ks[0][0], max(k[1] for k in ks)
busy, cur_s, cur_e = 0.0, None, None
gaps = []
for s, e, n in ks:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, last, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
    last = n
busy += cur_e - cur_s
tot = sum(e - s for s, e, _ in ks)
print(f"2 steps: span {(t1 - t0) / 1e3:.1f} ms, union busy {busy / 1e3:.1f} ms, sum of kernel times {tot / 1e3:.1f} ms, idle {(t1 - t0 - busy) / 1e3:.1f} ms over {len(gaps)} gaps")
import collections
by = collections.Counter()
for g, a, b in gaps:
    by[(a[:40], b[:40])] += g
for (a, b), g in by.most_common(25):
    print(f"  {g / 1e3:7.2f} ms idle between  {a}  ->  {b}")
