"""Host-only: print the tap-GEMM planner's choice (sos_conv2d_plan) for the benchmark layers, forward and data gradient.
usage: python scripts/plan_dump.py [B]   (runs without a GPU)"""
import ctypes as C
import sys
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T = 203
LAYERS = [
    ("x96_k5_d1", 96, 96, (5, 5), (1, 1), 256, T), ("x96_k5_d32x32", 96, 96, (5, 5), (32, 32), 256, T), ("x96_k7x1", 96, 96, (7, 1), (1, 1), 256, T),
    ("x2to96_k1x7", 16, 96, (1, 7), (1, 1), 256, T), ("x96to8_k1", 96, 8, (1, 1), (1, 1), 256, T),
    ("n48_k5_d1", 48, 48, (5, 5), (1, 1), 256, T), ("n48_k5_d4x4", 48, 48, (5, 5), (4, 4), 256, T),
    ("n48_k5_d16x1", 48, 48, (5, 5), (16, 1), 256, T), ("n48_k5_d32x32", 48, 48, (5, 5), (32, 32), 256, T),
    ("in_128_k5v", 128, 128, (5, 5), (1, 1), 132, 106), ("in_256_k3v", 256, 256, (3, 3), (1, 1), 66, 53),
]
names = "fast_is_w share+2wide+4merge g S groups stages stage_B grid cbe chunks N ec FB SB n_stg smem".split()
lib = _lib.lib()
for name, Cin, Cout, k, d, H, W in LAYERS:
    valid = name.endswith("v")
    offs = [((a - (0 if valid else (k[0] - 1) // 2)) * d[0], (b - (0 if valid else (k[1] - 1) // 2)) * d[1]) for a in range(k[0]) for b in range(k[1])]
    OH, OW = (H - (k[0] - 1) * d[0], W - (k[1] - 1) * d[1]) if valid else (H, W)
    a = _lib.ConvArgs()
    dh = (C.c_int32 * len(offs))(*[o[0] for o in offs]); dw = (C.c_int32 * len(offs))(*[o[1] for o in offs])
    a.tap_dh, a.tap_dw = dh, dw
    a.N, a.H, a.W, a.Cin, a.Cout, a.OH, a.OW, a.ntaps, a.stride = B, H, W, Cin, Cout, OH, OW, len(offs), 1
    a.YH, a.YW, a.Cy, a.osh, a.osw = OH, OW, (Cout + 7) // 8 * 8, 1, 1
    a.force_plan, a.x_dtype, a.y_dtype = -1, 1, 1
    info = (C.c_int32 * 16)()
    rc = lib.sos_conv2d_plan(C.byref(a), info)
    print(f"{name:16s}", "rc", rc, " ".join(f"{n}={v}" for n, v in zip(names, info)))

print("-- small-channel layers")
def q(name, Cin, Cout, offs, H, W, OH, OW, ydt=0):
    a = _lib.ConvArgs()
    dh = (C.c_int32 * len(offs))(*[o[0] for o in offs]); dw = (C.c_int32 * len(offs))(*[o[1] for o in offs])
    a.tap_dh, a.tap_dw = dh, dw
    a.N, a.H, a.W, a.Cin, a.Cout, a.OH, a.OW, a.ntaps, a.stride = B, H, W, Cin, Cout, OH, OW, len(offs), 1
    a.YH, a.YW, a.Cy, a.osh, a.osw = OH, OW, (Cout + 7) // 8 * 8, 1, 1
    a.force_plan, a.x_dtype, a.y_dtype = -1, 1, ydt
    info = (C.c_int32 * 16)()
    rc = lib.sos_conv2d_plan(C.byref(a), info)
    print(f"{name:16s}", "rc", rc, " ".join(f"{n}={v}" for n, v in zip(names, info)))
k3 = [(a, b) for a in range(3) for b in range(3)]
q("in_64to2 fwd", 64, 2, k3, 258, 205, 256, 203, 1)
q("in_64to2 dgrad", 16, 64, [(-a, -b) for a, b in k3], 256, 203, 258, 205, 0)
k5 = [(a, b) for a in range(5) for b in range(5)]
q("in_2to64 fwd", 16, 64, k5, 260, 207, 256, 203, 1)
q("in_2to64 dgrad", 64, 2, [(-a, -b) for a, b in k5], 256, 203, 260, 207, 0)
q("x2to96 fwd", 16, 96, [(0, b - 3) for b in range(7)], 256, 203, 256, 203, 1)
q("x2to96 dgrad", 96, 2, [(0, 3 - b) for b in range(7)], 256, 203, 256, 203, 0)
q("x96to8 fwd", 96, 8, [(0, 0)], 256, 203, 256, 203, 1)
q("x96to8 dgrad", 8, 96, [(0, 0)], 256, 203, 256, 203, 1)
