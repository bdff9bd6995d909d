import sys, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import ops, transform
ops.init()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
w = torch.randn(B, 32000, device="cuda") * 0.1
for _ in range(3):
    transform.stft_batch(w)
torch.cuda.synchronize()
