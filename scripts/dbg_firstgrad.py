"""Why do first-layer weight gradients deviate?  Records (x, dy) of every wgrad call, recomputes the weight gradient in
fp32 (and fp64) with plain torch on the GPU from the SAME operands, and compares with the kernel's result and the golden."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import ops, networks, layers as L
from oracle import nets
ops.init()
dev = torch.device("cuda:0")
gold = np.load("tests/golden/nets.npz")
rec = []
orig = L._conv_wgrad
def spy(x, dy, w, g):
    out = orig(x, dy, w, g)
    rec.append((x, dy, w, g, out))
    return out
L._conv_wgrad = spy

def ref_wgrad(x, dy, w, g, dtype):
    x, dy = x.to(dtype), dy.to(dtype)
    Cout, Cin = w.shape[0], w.shape[1]
    N, H, W, _ = x.shape
    OH, OW = dy.shape[1], dy.shape[2]
    out = torch.zeros(Cout, Cin, g.kh, g.kw, dtype=dtype, device=x.device)
    for (a, b), (oh, ow) in zip(g.taps, g.off):
        # rows oh..oh+OH of x (zero outside)
        xs = torch.zeros(N, OH, OW, x.shape[3], dtype=dtype, device=x.device)
        h0, h1 = max(0, -oh), min(OH, H - oh)
        w0, w1 = max(0, -ow), min(OW, W - ow)
        xs[:, h0:h1, w0:w1] = x[:, h0 + oh:h1 + oh, w0 + ow:w1 + ow]
        out[:, :, a, b] = torch.einsum("nhwo,nhwi->oi", dy[..., :Cout], xs[..., :Cin])
    return out

for mode in ("eval", "train"):
    rec.clear()
    sid = networks.get_network()
    sid.load_state_dict(nets.synth_state_dict(nets.sid_shapes(), 3))
    sid = sid.to(dev).train(mode == "train")
    x, lab = torch.tensor(gold["x"], device=dev), torch.tensor(gold["label"], device=dev)
    logits = sid(x, lab.shape[1])
    loss = L.BCEWithLogitsLoss.apply(logits, lab)
    loss.backward()
    torch.cuda.synchronize()
    print(f"--- SID {mode}: {len(rec)} wgrad calls (last = first layer)")
    for i, (xx, dy, w, g, out) in enumerate(rec):
        if g.stride != 1 or g.kind != "zero":
            continue
        r32 = ref_wgrad(xx, dy, w, g, torch.float32)
        r64 = ref_wgrad(xx, dy, w, g, torch.float64)
        sc = float(r64.abs().max())
        e_k = float((out.double() - r64).abs().max()) / sc
        e_32 = float((r32.double() - r64).abs().max()) / sc
        # cancellation measure: sum |terms| vs |sum|
        print(f"  call {i:2d} Cin {w.shape[1]:3d} Cout {w.shape[0]:3d} k{g.kh}x{g.kw} d{g.dh}x{g.dw}: kernel-vs-fp64 {e_k:.2e}  torchfp32-vs-fp64 {e_32:.2e}  |dw|max {sc:.3e} |x|max {float(xx.abs().max()):.2e} |dy|max {float(dy.abs().max()):.2e}")
    name = "encoder_audio.0.block.0.weight"
    want = gold[f"sid_{mode}_grad:{name}"]
    got = dict(sid.named_parameters())[name].grad.cpu().numpy()
    xx, dy, w, g, out = rec[-1]
    r64 = ref_wgrad(xx, dy, w, g, torch.float64).cpu().numpy()
    sc = np.abs(want).max()
    print(f"  first layer: kernel-vs-golden {np.abs(got-want).max()/sc:.2e}; fp64(our dy,x)-vs-golden {np.abs(r64-want).max()/sc:.2e}")
