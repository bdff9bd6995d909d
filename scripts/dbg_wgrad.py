"""Structured-input probe of sos_conv2d_wgrad (prints what the tensor core actually computed)."""
import sys, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import ops
ops.init()
dev = torch.device("cuda:0")
torch.set_printoptions(linewidth=220, precision=1, sci_mode=False)
N, H, W, Cin, Cout = 1, 16, 24, 32, 32
P = N * H * W
def run(x, dy, tag, **kw):
    info = [0] * 8
    dw = ops.conv_wgrad(x.contiguous(), dy.contiguous(), [0], [0], Cout, H, W, 1, plan_out=info, **kw)
    torch.cuda.synchronize()
    ref = dy.reshape(P, Cout).t() @ x.reshape(P, Cin)
    err = float((dw[0] - ref).abs().max() / (ref.abs().max() + 1e-9))
    print(f"== {tag}: plan {info} err {err:.3e}")
    if err > 1e-2:
        print("got[:8,:12]\n", dw[0][:8, :12].cpu())
        print("ref[:8,:12]\n", ref[:8, :12].cpu())
    return dw
ones_x = torch.ones(N, H, W, Cin, device=dev)
ones_dy = torch.ones(N, H, W, Cout, device=dev)
run(ones_x, ones_dy, "ones*ones")
ci = torch.arange(Cin, device=dev, dtype=torch.float32).expand(N, H, W, Cin)
co = torch.arange(Cout, device=dev, dtype=torch.float32).expand(N, H, W, Cout)
run(ci, ones_dy, "x=ci")
run(ones_x, co, "dy=co")
pix = torch.arange(P, device=dev, dtype=torch.float32).view(N, H, W, 1)
onehot = (pix == 5).float()
run(ci * onehot, co, "pixel 5 only")
run(torch.randn(N, H, W, Cin, device=dev), torch.randn(N, H, W, Cout, device=dev), "random")
for plan in (0, 1, 2, 3):
    try:
        run(torch.randn(N, H, W, Cin, device=dev), torch.randn(N, H, W, Cout, device=dev), f"random plan {plan}", force_plan=plan)
    except Exception as e:
        print("plan", plan, "->", e)
