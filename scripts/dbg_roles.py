"""Per-role wait cycles of the tap GEMM (run with SOS_EPI_DBG=16): light (96 -> 8 1x1, 64 -> 64 1x1, 64 -> 2 3x3), mid (96 -> 96 7x1) and
heavy (96 -> 96 5x5) layers.  Each run() prints the counters of the warm-up forward call first, then those of the named role."""
import os, sys, torch
sys.path.insert(0, '.')
import sos_b200
from sos_b200 import layers as L, ops
ops.init()
dev = torch.device('cuda')
B, T = 32, 203
def run(name, kind, Cin, Cout, k, d, H, W, role):
    g = L.ConvGeom(kind, k[0], k[1], d[0], d[1], 1)
    x = ops.to_half(torch.randn(B, H, W, Cin, device=dev))
    w = torch.randn(Cout, Cin, k[0], k[1], device=dev) * 0.05
    OH, OW = g.out_size(H, W)
    y = L._conv_forward(x, w, g)
    dy = ops.to_half(torch.randn_like(y))
    torch.cuda.synchronize()
    print("==", name, role, flush=True)
    if role == "fwd":
        L._conv_forward(x, w, g, want_stats=True, y_half=True)
    else:
        L._conv_dgrad_raw(dy, w, g, x.shape, None, y_half=(kind == "zero"))
    torch.cuda.synchronize()
run("x96to8_k1", "zero", 96, 8, (1, 1), (1, 1), 256, T, "dgrad")
run("x96to8_k1", "zero", 96, 8, (1, 1), (1, 1), 256, T, "fwd")
run("k1_64to64", "zero", 64, 64, (1, 1), (1, 1), 256, T, "fwd")
run("in_64to2_k3", "valid", 64, 2, (3, 3), (1, 1), 258, T + 2, "dgrad")
run("x96_k5", "zero", 96, 96, (5, 5), (1, 1), 256, T, "fwd")
run("x96_k7x1", "zero", 96, 96, (7, 1), (1, 1), 256, T, "fwd")
