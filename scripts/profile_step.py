"""One training step (BASELINE configs[1]) inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`.
usage: python scripts/profile_step.py [B]"""
import sys
import torch
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import agent as ag, ops, tools, transform
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
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


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
