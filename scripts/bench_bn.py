"""BatchNorm backward (reduce + finalize + apply) of the half path at the two big map shapes; GB/s of the algorithmic bytes."""
import sys, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import layers as L, ops
ops.init(); dev = torch.device("cuda:0")
N, H, W = 32, 256, 203
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
for C in (48, 96):
    dz = ops.to_half(torch.randn(N, H, W, C, device=dev) * 0.5)
    yb = ops.to_half(torch.randn(N, H, W, C, device=dev))
    stats = [torch.randn(C, device=dev) * 0.1, torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) + 0.5, torch.randn(C, device=dev) * 0.1]
    inv = torch.ones(1, device=dev)
    part = torch.zeros(148, 4, C, device=dev)
    full = t(lambda: ops.bn_train_backward_half(dz, yb, stats, ops.ACT_RELU, None, dz_inv=inv))
    pre = t(lambda: ops.bn_train_backward_half(dz, yb, stats, ops.ACT_RELU, None, dz_inv=inv, pre_partial=part))
    el = N * H * W * C
    print(f"C={C}: backward {full*1e3:.1f} us, without the reduce pass {pre*1e3:.1f} us -> reduce {1e3*(full-pre):.1f} us = {4*el/(full-pre)/1e6:.0f} GB/s; apply+finalize {6*el/pre/1e6:.0f} GB/s")
