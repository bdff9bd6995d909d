"""The small-channel layers (2->64, 2->96, 96->8, 64->2 ...): time and planner choice per role and per forced plan."""
import sys
import torch
sys.path.insert(0, ".")
import sos_b200  # noqa
from sos_b200 import layers as L, ops

ops.init()
dev = torch.device("cuda:0")
B, T = 32, 203


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
    return e0.elapsed_time(e1) / n * 1e3


CASES = [("in_64to2_k3", "valid", 64, 2, (3, 3), 258, T + 2), ("x96to8_k1", "zero", 96, 8, (1, 1), 256, T), ("n48to4_k1", "zero", 48, 4, (1, 1), 256, T),
         ("in_2to64_k5", "valid", 16, 64, (5, 5), 260, T + 4), ("x2to96_k1x7", "zero", 16, 96, (1, 7), 256, T)]
for name, kind, Cin, Cout, k, H, W in CASES:
    g = L.ConvGeom(kind, k[0], k[1], 1, 1, 1)
    x = ops.to_half(torch.randn(B, H, W, Cin, device=dev).relu_())
    w = torch.randn(Cout, 2 if Cin == 16 else Cin, k[0], k[1], device=dev) * 0.05
    OH, OW = g.out_size(H, W)
    y = L._conv_forward(x, w, g)
    dy = ops.to_half(torch.randn_like(y))
    one = torch.ones(1, device=dev)
    print(f"{name}: y {tuple(y.shape)}")
    # forward / dgrad with every forced plan
    for role in ("fwd", "dgrad"):
        for plan in (-1, 0, 1, 2, 3):
            info = [0] * 8
            try:
                if role == "fwd":
                    wk = L._pack_fwd(w, g.taps, Cin, True)
                    fn = lambda: ops.conv_tc(x, wk, [o[0] for o in g.off], [o[1] for o in g.off], Cout, OH, OW, 1, force_plan=plan, plan_out=info)
                else:
                    wk = L._pack_fwd(w.permute(1, 0, 2, 3), g.taps, dy.shape[3], True)
                    fn = lambda: ops.conv_tc(dy, wk, [-o[0] for o in g.off], [-o[1] for o in g.off], w.shape[1], H, W, 1, force_plan=plan,
                                             plan_out=info, out_scale=one)
                t = timeit(fn)
                print(f"   {role:5s} plan {plan:2d}: {t:7.1f} us  fast_is_w {info[0]} share {info[1]} g {info[2]} S {info[3]} groups {info[4]} stages {info[5]} stage_bytes {info[6]}")
            except Exception as e:
                print(f"   {role:5s} plan {plan:2d}: {str(e)[:80]}")
    info = [0] * 8
    t = timeit(lambda: ops.conv_wgrad(x, dy, [o[0] for o in g.off], [o[1] for o in g.off], dy.shape[3], OH, OW, 1, plan_out=info))
    print(f"   wgrad        : {t:7.1f} us  plan {info}")
