"""Per-layer weight-gradient deviation of the CUDA path against the functional oracle run in fp32 ON THE GPU (cuDNN, TF32 off)
and, for scale, against the same oracle with cuDNN TF32 on."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
import sos_b200
from sos_b200 import ops, networks, layers as L, transform
from oracle import nets
ops.init()
dev = torch.device("cuda:0")
gold = np.load("tests/golden/nets.npz")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

def oracle_grads(kind, mode, tf32):
    torch.backends.cudnn.allow_tf32 = tf32
    x, lab = torch.tensor(gold["x"], device=dev), torch.tensor(gold["label"], device=dev)
    if kind == "sid":
        sd = {k: v.to(dev) for k, v in nets.synth_state_dict(nets.sid_shapes(), 3).items()}
        for k, v in sd.items():
            if v.is_floating_point() and 'running' not in k: v.requires_grad_(True)
        logits = nets.sid_forward(sd, x, lab.shape[1], training=(mode == "train"))
        loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, lab)
    else:
        sd = {k: v.to(dev) for k, v in nets.synth_state_dict(nets.joint_shapes(), 4).items()}
        for k, v in sd.items():
            if v.is_floating_point() and 'running' not in k: v.requires_grad_(True)
        n = torch.tensor(gold["n"], device=dev)
        n_pred, mask = nets.joint_forward(sd, x, n, training=(mode == "train"))
        from oracle import transform as otf
        rec = otf.batch_fast_icRM_sigmoid(x, mask)
        loss = torch.nn.functional.mse_loss(n_pred, torch.tensor(gold["tgt_n"], device=dev)) + \
            torch.nn.functional.mse_loss(rec, torch.tensor(gold["tgt_c"], device=dev))
    loss.backward()
    torch.backends.cudnn.allow_tf32 = False
    return {k: v.grad for k, v in sd.items() if v.is_floating_point() and v.grad is not None}

def ours(kind, mode):
    x, lab = torch.tensor(gold["x"], device=dev), torch.tensor(gold["label"], device=dev)
    if kind == "sid":
        net = networks.get_network()
        net.load_state_dict(nets.synth_state_dict(nets.sid_shapes(), 3))
        net = net.to(dev).train(mode == "train")
        loss = L.BCEWithLogitsLoss.apply(net(x, lab.shape[1]), lab)
    else:
        net = networks.get_network(object())
        net.load_state_dict(nets.synth_state_dict(nets.joint_shapes(), 4))
        net = net.to(dev).train(mode == "train")
        n = torch.tensor(gold["n"], device=dev)
        n_pred, mask = net(x, n)
        rec = transform.batch_fast_icRM_sigmoid(x, mask)
        loss = L.MSELoss.apply(n_pred, torch.tensor(gold["tgt_n"], device=dev)) + L.MSELoss.apply(rec, torch.tensor(gold["tgt_c"], device=dev))
    loss.backward()
    return {k: v.grad for k, v in net.named_parameters()}

for kind in ("sid", "joint"):
    for mode in ("train", "eval"):
        try:
            ref = oracle_grads(kind, mode, False)
            tf = oracle_grads(kind, mode, True)
        except Exception as e:
            print(kind, mode, "oracle failed:", repr(e)); continue
        got = ours(kind, mode)
        print(f"=== {kind} {mode}: relative max error of weight gradients (ours | cuDNN-TF32 oracle) vs fp32 oracle")
        for k in ref:
            if k.endswith(".weight") and ref[k].dim() == 4:
                sc = float(ref[k].abs().max()) + 1e-20
                print(f"  {k:38s} ours {float((got[k]-ref[k]).abs().max())/sc:.2e}   cudnn-tf32 {float((tf[k]-ref[k]).abs().max())/sc:.2e}")
