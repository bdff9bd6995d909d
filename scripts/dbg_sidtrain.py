"""Stage-by-stage comparison of the SID train-mode path with the functional oracle (fp32 on the GPU)."""
import sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import ops, networks, layers as L
from oracle import nets
ops.init()
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
gold = np.load("tests/golden/nets.npz")
x, lab = torch.tensor(gold["x"], device=dev), torch.tensor(gold["label"], device=dev)
V = lab.shape[1]
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-20))
for mode in ("train", "eval"):
    tr = mode == "train"
    sd = {k: v.to(dev) for k, v in nets.synth_state_dict(nets.sid_shapes(), 3).items()}
    for k, v in sd.items():
        if v.is_floating_point() and "running" not in k: v.requires_grad_(True)
    f_o = nets._encoder(sd, "encoder_audio", x, nets.SID_KS, nets.SID_DL, tr, None); f_o.retain_grad()
    seq_o = F.interpolate(f_o.reshape(f_o.size(0), -1, f_o.size(3)), size=V).permute(2, 0, 1); seq_o.retain_grad()
    m_o = nets._lstm(sd, "lstm", seq_o, 100); m_o.retain_grad()
    h_o = F.relu(m_o.permute(1, 0, 2) @ sd["fc1.0.weight"].t() + sd["fc1.0.bias"])
    lo_o = (h_o @ sd["fc1.2.weight"].t() + sd["fc1.2.bias"]).squeeze(2); lo_o.retain_grad()
    F.binary_cross_entropy_with_logits(lo_o, lab).backward()

    sid = networks.get_network(); sid.load_state_dict(nets.synth_state_dict(nets.sid_shapes(), 3)); sid = sid.to(dev).train(tr)
    f = sid.encoder_audio(networks._nhwc_in(x)); f.retain_grad()
    seq = L.FeatToSeq.apply(V, (8,), f); seq.retain_grad()
    m = sid.lstm(seq); m.retain_grad()
    lo = sid.fc1(m.permute(1, 0, 2)).squeeze(2); lo.retain_grad()
    L.BCEWithLogitsLoss.apply(lo, lab).backward()
    f_n = f.permute(0, 3, 1, 2)
    print(f"--- {mode}")
    print("  encoder out  ", rel(f_n, f_o.detach()), " |f| max", float(f_o.abs().max()))
    print("  seq          ", rel(seq, seq_o.detach()))
    print("  lstm out     ", rel(m, m_o.detach()))
    print("  logits       ", rel(lo, lo_o.detach()), float((lo - lo_o).abs().max()))
    print("  d logits     ", rel(lo.grad, lo_o.grad))
    print("  d lstm out   ", rel(m.grad, m_o.grad))
    print("  d seq        ", rel(seq.grad, seq_o.grad))
    print("  d encoder out", rel(f.grad.permute(0, 3, 1, 2), f_o.grad))
    ps = dict(sid.named_parameters())
    for k in ("fc1.2.weight", "fc1.2.bias", "fc1.0.weight", "fc1.0.bias", "lstm.bias_ih_l0", "lstm.weight_hh_l0", "lstm.weight_ih_l0", "lstm.bias_ih_l0_reverse",
              "encoder_audio.11.block.1.weight", "encoder_audio.11.block.1.bias", "encoder_audio.11.block.0.weight"):
        print(f"  grad {k:34s}", rel(ps[k].grad, sd[k].grad))
