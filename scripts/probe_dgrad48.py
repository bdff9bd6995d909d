"""Why is the in-step 48-channel data gradient 1.7x slower than the forward of the same shape?  Times the same tap GEMM with
different operand CONTENTS and with / without the output scale."""
import sys
import torch
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import layers as L, ops

ops.init()
dev = torch.device("cuda:0")
B, H, W = 32, 256, 203


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for C in (48, 96):
    g = L.ConvGeom("zero", 5, 5, 1, 1, 1)
    w = torch.randn(C, C, 5, 5, device=dev) * 0.05
    base = torch.randn(B, H, W, C, device=dev)
    one = torch.ones(1, device=dev)
    cases = {
        "randn": base,
        "relu(randn) 50% zeros": base.relu(),
        "randn * 1e-3": base * 1e-3,
        "randn * 30": base * 30,
        "zeros": base * 0,
        "const 1": base * 0 + 1,
        "uniform 0..1": torch.rand_like(base),
    }
    for name, t in cases.items():
        a = ops.to_half(t.contiguous())
        t_f = timeit(lambda: L._conv_forward(a, w, g))
        t_d = timeit(lambda: L._conv_dgrad(a, w, g, a.shape))
        t_s = timeit(lambda: L._conv_dgrad(a, w, g, a.shape, one))
        print(f"C={C} {name:24s} fwd {t_f*1e3:7.1f} us   dgrad {t_d*1e3:7.1f} us   dgrad+out_scale {t_s*1e3:7.1f} us")
