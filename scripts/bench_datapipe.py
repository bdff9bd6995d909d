"""Device-side training-item construction (datapipe.make_joint_items) for a 32-clip batch: time and algorithmic HBM bytes."""
import sys
import torch
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import datapipe, ops

ops.init()
dev = torch.device("cuda:0")
B, L = int(sys.argv[1]) if len(sys.argv) > 1 else 32, 32000
g = torch.Generator(device="cpu").manual_seed(0)
audio = (torch.randn(B, L, generator=g) * 0.1).to(dev)
noise = (torch.randn(B, L, generator=g) * 0.3).to(dev)
bits = (torch.rand(B, 60, generator=g) > 0.4).to(torch.uint8).to(dev)
snr = torch.tensor([datapipe.SNRS[i % 7] for i in range(B)], dtype=torch.float32, device=dev)
for _ in range(3):
    datapipe.make_joint_items(audio, noise, snr, bits)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
n = 20
for _ in range(n):
    item = datapipe.make_joint_items(audio, noise, snr, bits)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
T = item["mixed"].shape[3]
# waveforms: gate (r+w), add_signals (2r x3 passes from L2, 3w), gate (r+w); STFT 4 signals (r wave, w spec); cRM (2r + w)
alg = 4.0 * B * L * (2 + 2 + 3 + 2 + 4) + 4.0 * B * 2 * 256 * T * (4 + 3)
print(f"make_joint_items B={B}: {ms * 1e3:.1f} us per batch  ({B / ms * 1e3:.0f} clips/s), {alg / 1e6:.1f} MB algorithmic -> {alg / ms / 1e6:.0f} GB/s")
