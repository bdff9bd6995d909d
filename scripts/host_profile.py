"""Where the HOST time of one training step goes (cProfile over 3 steps, no device synchronisation inside)."""
import cProfile
import pstats
import sys
import torch
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


for _ in range(3):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)
