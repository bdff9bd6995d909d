"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (this container
only: /root/reference does not exist on the GPU box).  TEST INFRASTRUCTURE.

  python -m oracle.make_golden

* networks: M1/networks.py, M2/networks.py are imported by path (torch only).
* transform.py needs `import librosa` (third-party, librosa==0.7.1, absent):
  a stub module is injected whose stft/istft are oracle.transform's numpy
  restatement, so the fixture pins everything in transform.py EXCEPT librosa's
  own FFT code (that leg is cross-checked against torch.stft in the tests).
* tools.py imports matplotlib/PIL/...; the two pure-numpy functions on the hot
  path are extracted from the file with `ast` and exec'd unmodified.
"""
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
M1 = REF + "/model_1_silent_interval_detection/audioonly_model"
M2 = REF + "/model_2_audio_denoising/audio_denoising_model"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def _extract(path, names, ns):
    tree = ast.parse(open(path).read())
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def main():
    from oracle import nets, transform as otf, synth
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(8)

    # ---- networks --------------------------------------------------------
    m1 = _load(M1 + "/networks.py", "ref_m1_networks")
    m2 = _load(M2 + "/networks.py", "ref_m2_networks")

    class Cfg:
        kernel_sizes, dilations = nets.KS, nets.DL

    B, T, V = 2, 71, 21
    g = torch.Generator().manual_seed(77)
    x = 0.5 * torch.randn(B, 2, 256, T, generator=g)
    n = x * (torch.rand(B, 1, 1, T, generator=g) > 0.5)
    tgt_n = 0.5 * torch.randn(B, 2, 256, T, generator=g)
    tgt_c = 0.5 * torch.randn(B, 2, 256, T, generator=g)
    lab = (torch.rand(B, V, generator=g) > 0.5).float()
    out = {"x": x.numpy(), "n": n.numpy(), "tgt_n": tgt_n.numpy(), "tgt_c": tgt_c.numpy(), "label": lab.numpy()}

    sid = m1.get_network()
    sid.load_state_dict(nets.synth_state_dict(nets.sid_shapes(), 3))
    for mode in ("eval", "train"):
        sid.train(mode == "train")
        sid.zero_grad()
        logits = sid(x, V)
        loss = torch.nn.BCEWithLogitsLoss()(logits, lab)
        loss.backward()
        out[f"sid_{mode}_logits"] = logits.detach().numpy()
        out[f"sid_{mode}_loss"] = np.float32(loss.item())
        for k in ("encoder_audio.0.block.0.weight", "encoder_audio.5.block.1.weight", "encoder_audio.11.block.0.weight",
                  "lstm.bias_ih_l0", "fc1.0.bias"):
            out[f"sid_{mode}_grad:{k}"] = dict(sid.named_parameters())[k].grad.numpy().copy()
    out["sid_train_rm:encoder_audio.3.block.1.running_mean"] = sid.state_dict()["encoder_audio.3.block.1.running_mean"].numpy().copy()
    out["sid_train_rv:encoder_audio.3.block.1.running_var"] = sid.state_dict()["encoder_audio.3.block.1.running_var"].numpy().copy()

    sys.modules["librosa"] = types.SimpleNamespace(
        stft=lambda y, n_fft, hop, win: otf.librosa_stft(y, n_fft, hop, win),
        istft=lambda S, hop, win: otf.librosa_istft(S, hop, win))
    tr = _load(M2 + "/transform.py", "ref_m2_transform")

    for spread in (False, True):
        tag = "spread" if spread else "plain"
        joint = m2.get_network(Cfg())
        joint.load_state_dict(nets.synth_state_dict(nets.joint_shapes(), 4, spread=spread))
        for mode in ("eval", "train"):
            joint.train(mode == "train")
            joint.zero_grad()
            n_pred, mask = joint(x, n)
            rec = tr.batch_fast_icRM_sigmoid(x, mask)
            l1 = torch.nn.MSELoss()(n_pred, tgt_n)
            l2 = torch.nn.MSELoss()(rec, tgt_c)
            (l1 + l2).backward()
            pre = f"joint_{tag}_{mode}"
            out[pre + "_npred"] = n_pred.detach().numpy()
            out[pre + "_mask"] = mask.detach().numpy()
            out[pre + "_loss1"] = np.float32(l1.item())
            out[pre + "_loss2"] = np.float32(l2.item())
            if not spread:
                params = dict(joint.named_parameters())
                for k in ("stage1.down1.0.block.1.weight", "stage1.mid.4.block.2.weight", "stage1.mid.8.block.0.weight",
                          "stage1.up2.1.block.1.bias", "stage1.mid.3.block.3.weight",
                          "stage2.encoder_x.0.block.0.weight", "stage2.encoder_x.9.block.1.bias",
                          "stage2.encoder_n.14.block.0.weight", "stage2.lstm.bias_hh_l0_reverse", "stage2.fc.4.bias"):
                    gr = params[k].grad.numpy()
                    out[f"{pre}_grad:{k}"] = (gr[:16, :16] if gr.ndim == 4 and gr.size > 100000 else gr).copy()
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **out)

    # ---- transform.py (reference functions, stub librosa) ------------------
    clip = synth.make_clip(5, length=6000)
    mixed, clean = clip["mixed"], clip["clean"]
    Fm, Fc = tr.fast_stft(mixed), tr.fast_stft(clean)
    crm = tr.fast_cRM_sigmoid(Fc, Fm)
    t_out = {
        "mixed": mixed, "clean": clean,
        "fast_stft_mixed": Fm.astype(np.float32),
        "fast_cRM_sigmoid": crm,
        "fast_icRM_sigmoid": tr.fast_icRM_sigmoid(Fm, crm),
        "fast_istft_mixed": tr.fast_istft(Fm).astype(np.float32),
        "generate_cRM": tr.generate_cRM(Fm, Fc),
        "cRM_sigmoid_recover": tr.cRM_sigmoid_recover(crm),
    }
    Yb = torch.tensor(Fm.transpose(2, 0, 1)[None], dtype=torch.float32)
    Cb = torch.tensor(crm.transpose(2, 0, 1)[None], dtype=torch.float32)
    t_out["batch_fast_icRM_sigmoid"] = tr.batch_fast_icRM_sigmoid(Yb, Cb).numpy()
    np.savez_compressed(os.path.join(OUT, "transform.npz"), **t_out)

    # ---- tools.py: bit mask + add_signals (reference source exec'd) --------
    from itertools import groupby
    ns = {"np": np, "groupby": groupby}
    _extract(M2 + "/tools.py", {"convert_bitstreammask_to_audiomask", "add_signals", "power_of_signal"}, ns)
    rng = np.random.default_rng(11)
    g_out = {}
    cases = [(28000, 14000 / 30, 60), (32000, 16000 / 30, 60), (32000, 16000 / 30, 59), (31990, 16000 / 30, 60),
             (7000, 14000 / 25, 13), (160000, 16000 / 30, 300), (32000, 16000 / 30, 61)]
    for ci, (length, ratio, nb) in enumerate(cases):
        for rep in range(3):
            bits = "".join(rng.choice(["0", "1"], p=[0.5, 0.5]) for _ in range(nb)) if rep < 2 else ("0" * nb if ci % 2 else "01" * (nb // 2) + "0" * (nb % 2))
            mask = ns["convert_bitstreammask_to_audiomask"](np.zeros(length, dtype=np.float32), ratio, bits)
            g_out[f"mask:{length}:{ratio!r}:{bits}"] = np.packbits(mask.astype(np.uint8))
    sig = rng.standard_normal(4000) * (rng.random(4000) > 0.3)
    noi = rng.standard_normal(4000)
    for snr in (-10, 0, 7):
        mx, cl, ns_ = ns["add_signals"](sig, [noi], snr, norm=0.5)
        g_out[f"add_signals:{snr}"] = np.stack([mx, cl, ns_[0]])
    g_out["add_signals:sig"], g_out["add_signals:noise"] = sig, noi
    np.savez_compressed(os.path.join(OUT, "tools.npz"), **g_out)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
