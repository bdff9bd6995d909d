"""Per-role wait cycles of the weight-gradient kernel (SOS_WGRAD_DBG=1) on the 48- and 96-channel 5x5 layers."""
import sys, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import layers as L, ops
ops.init(); dev = torch.device("cuda")
for C in (48, 96):
    g = L.ConvGeom("zero", 5, 5, 1, 1, 1)
    x = ops.to_half(torch.randn(32, 256, 203, C, device=dev))
    dy = ops.to_half(torch.randn(32, 256, 203, C, device=dev))
    w = torch.randn(C, C, 5, 5, device=dev)
    torch.cuda.synchronize()
    print("==", C, flush=True)
    L._conv_wgrad_raw(x, dy, w, g, torch.ones(1, device=dev))
    torch.cuda.synchronize()
