"""Accumulation accuracy of the kind::f16 vs kind::tf32 tap GEMM on IDENTICAL operand values (half-representable, so both
kernels multiply exactly the same numbers): error vs a float64 convolution, relative to the output scale."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import layers as L, ops

ops.init()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
for Cin, Cout, k in ((96, 96, 5), (48, 48, 5), (256, 256, 3)):
    for kind in ("randn", "int"):
        N, H, W = 2, 32, 40
        if kind == "randn":
            x = torch.randn(N, Cin, H, W, generator=g).half().float()
            w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).half().float()
        else:
            x = torch.randint(-8, 9, (N, Cin, H, W), generator=g).float()
            w = torch.randint(-4, 5, (Cout, Cin, k, k), generator=g).float()
        ref = F.conv2d(x.double(), w.double(), None, 1, k // 2).float()
        geom = L.ConvGeom("zero", k, k, 1, 1, 1)
        x32 = ops.nchw_to_nhwc(x.to(dev), Cin)
        y32 = L._conv_forward(x32, w.to(dev), geom)
        xh = ops.to_half(x32)
        yh = L._conv_forward(xh, w.to(dev), geom)
        torch.cuda.synchronize()
        sc = float(ref.abs().max())
        e32 = float((ops.nhwc_to_nchw(y32, Cout).cpu() - ref).abs().max()) / sc
        eh = float((ops.nhwc_to_nchw(yh, Cout).cpu() - ref).abs().max()) / sc
        print(f"{Cin}->{Cout} k{k} {kind}: scale {sc:.3g}  tf32 kernel err {e32:.2e}  f16 kernel err {eh:.2e}")
