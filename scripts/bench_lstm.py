"""Recurrence kernels alone: us per time step, forward and backward, at the two shapes of the hot path.
SOS_LSTM_CLUSTER=0 selects the global-memory-barrier kernels (lstm.cu) instead of the cluster kernels (lstm_cluster.cu)."""
import sys
import torch
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import ops

ops.init()
dev = torch.device("cuda:0")
for T, B, H in ((203, 32, 200), (60, 32, 100), (203, 1, 200)):
    gx = torch.randn(T, B, 2, 4 * H, device=dev) * 0.5
    whh = torch.randn(2, 4 * H, H, device=dev) * 0.05
    out, gates, cell = ops.lstm_forward(gx, whh)
    dout = torch.randn_like(out)
    res = []
    for fn in (lambda: ops.lstm_forward(gx, whh), lambda: ops.lstm_backward(dout, whh, out, gates, cell)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    print(f"T={T} B={B} H={H}: forward {res[0] * 1e3:.0f} us ({res[0] * 1e3 / T:.2f} us/step)   backward {res[1] * 1e3:.0f} us ({res[1] * 1e3 / T:.2f} us/step)")
