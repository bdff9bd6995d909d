import sys, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import layers as L, ops
ops.init(); dev = torch.device("cuda:0")
N, C, H, W = 32, 48, 256, 203
dy = ops.to_half(torch.randn(N, H, W, C, device=dev) * 0.5)
w = torch.randn(C, C, 5, 5, device=dev) * 0.03
yb = ops.to_half(torch.randn(N, H, W, C, device=dev))
stats = [torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1]
g = L.ConvGeom("zero", 5, 5, 1, 1, 1)
def t(fn, n=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
print("dgrad plain %.3f ms" % t(lambda: L._conv_dgrad_raw(dy, w, g, (N, H, W, C), None, y_half=True)))
print("dgrad + fused reduction %.3f ms" % t(lambda: L._conv_dgrad_raw(dy, w, g, (N, H, W, C), None, y_half=True, bnr=(yb, stats))))
inv = torch.ones(1, device=dev)
dz, part = L._conv_dgrad_raw(dy, w, g, (N, H, W, C), None, y_half=True, bnr=(yb, stats))
print("bn backward full %.3f ms" % t(lambda: ops.bn_train_backward_half(dz, yb, stats, ops.ACT_RELU, None, dz_inv=inv)))
print("bn backward pre  %.3f ms" % t(lambda: ops.bn_train_backward_half(dz, yb, stats, ops.ACT_RELU, None, dz_inv=inv, pre_partial=part)))
