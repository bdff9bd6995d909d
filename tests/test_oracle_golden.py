"""CPU: the oracle (oracle/*.py) against the golden vectors generated FROM THE REFERENCE ITSELF (oracle/make_golden.py ran
the reference's own networks.py / transform.py / tools.py functions in the build container; see tests/golden/).
This is what pins the oracle; the GPU parity tests then compare the CUDA path with the oracle and with the same goldens."""
import os
import numpy as np
import pytest
import torch
import torch.nn.functional as F


@pytest.fixture(scope="module")
def g_tools(golden_dir):
    return np.load(golden_dir + "/tools.npz")


@pytest.fixture(scope="module")
def g_tr(golden_dir):
    return np.load(golden_dir + "/transform.npz")


@pytest.fixture(scope="module")
def g_nets(golden_dir):
    return np.load(golden_dir + "/nets.npz")


def test_bit_masks_bit_exact(g_tools):
    """convert_bitstreammask_to_audiomask (M2/tools.py:340-362): every fixture (ragged lengths, 25/30 fps, all-silent,
    alternating, bit strings one short / one long) must match bit for bit."""
    from oracle.gating import bits_to_sample_mask
    n = 0
    for key in g_tools.files:
        if not key.startswith("mask:"):
            continue
        _, length, ratio, bits = key.split(":")
        want = np.unpackbits(g_tools[key])[:int(length)]
        got = bits_to_sample_mask(int(length), float(ratio), bits).astype(np.uint8)
        assert np.array_equal(got, want), key[:60]
        n += 1
    assert n >= 20


def test_add_signals(g_tools):
    """add_signals (M2/tools.py:217-276) at three SNRs: mixed / clean / noise, float64."""
    from oracle import synth
    sig, noise = g_tools["add_signals:sig"], g_tools["add_signals:noise"]
    for snr in (-10, 0, 7):
        mixed, clean, noi = synth.add_signals(sig, noise, snr, norm=0.5)
        want = g_tools[f"add_signals:{snr}"]
        assert np.abs(mixed - want[0]).max() < 1e-12 and np.abs(clean - want[1]).max() < 1e-12 and np.abs(noi - want[2]).max() < 1e-12
        assert abs(np.abs(mixed).max() - 0.5) < 1e-12


def test_transform_functions(g_tr):
    """fast_stft / fast_istft / cRM compress + recover / batch_fast_icRM_sigmoid of M2/transform.py."""
    from oracle import transform as otf
    Fm, Fc = otf.fast_stft(g_tr["mixed"]), otf.fast_stft(g_tr["clean"])
    assert np.abs(Fm - g_tr["fast_stft_mixed"]).max() < 1e-6
    crm = otf.fast_cRM_sigmoid(Fc, Fm)
    assert np.abs(crm - g_tr["fast_cRM_sigmoid"]).max() < 1e-9
    assert np.abs(otf.generate_cRM(Fm, Fc) - g_tr["generate_cRM"]).max() < 1e-6 * np.abs(g_tr["generate_cRM"]).max()
    assert np.abs(otf.cRM_sigmoid_recover(g_tr["fast_cRM_sigmoid"]) - g_tr["cRM_sigmoid_recover"]).max() < 1e-9
    assert np.abs(otf.fast_icRM_sigmoid(Fm, g_tr["fast_cRM_sigmoid"]) - g_tr["fast_icRM_sigmoid"]).max() < 1e-6
    assert np.abs(otf.fast_istft(g_tr["fast_stft_mixed"]) - g_tr["fast_istft_mixed"]).max() < 1e-6
    Yb = torch.tensor(g_tr["fast_stft_mixed"].transpose(2, 0, 1)[None])
    Cb = torch.tensor(g_tr["fast_cRM_sigmoid"].transpose(2, 0, 1)[None], dtype=torch.float32)
    assert float((otf.batch_fast_icRM_sigmoid(Yb, Cb) - torch.tensor(g_tr["batch_fast_icRM_sigmoid"])).abs().max()) < 1e-4


def test_stft_restatement_against_torch(g_tr):
    """librosa 0.7.1 is absent: the numpy restatement of its stft/istft is cross-checked against
    torch.stft / torch.istft with the equivalent arguments, at the reference-native, benchmark and ragged lengths."""
    from oracle import transform as otf
    win = torch.hann_window(400, periodic=True)
    for L in (6000, 28000, 32000, 31990):
        y = np.random.default_rng(L).standard_normal(L).astype(np.float32) * 0.3
        S = otf.librosa_stft(y)
        ref = torch.stft(torch.tensor(y), 510, 158, 400, win, center=True, pad_mode="reflect", return_complex=True).numpy()
        assert S.shape == ref.shape == (256, 1 + L // 158)
        assert np.abs(S - ref).max() < 2e-5
        w = otf.librosa_istft(S)
        ref_w = torch.istft(torch.tensor(ref), 510, 158, 400, win, center=True, length=158 * (L // 158)).numpy()
        assert w.shape == (158 * (L // 158),)
        assert np.abs(w - ref_w).max() < 1e-5


def test_stft_edge_cases():
    from oracle import transform as otf
    assert otf.num_frames(32000) == 203 and otf.num_frames(28000) == 178 and otf.num_frames(160000) == 1013
    z = otf.librosa_stft(np.zeros(1000, np.float32))
    assert z.shape == (256, 7) and not z.any()
    # linearity and the istft(stft(x)) identity on the trimmed support
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal(4000).astype(np.float32), rng.standard_normal(4000).astype(np.float32)
    assert np.abs(otf.librosa_stft(a + 2 * b) - (otf.librosa_stft(a) + 2 * otf.librosa_stft(b))).max() < 1e-4
    rec = otf.librosa_istft(otf.librosa_stft(a))
    assert np.abs(rec - a[:rec.size]).max() < 1e-5


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_sid_oracle_matches_reference(g_nets, mode):
    """oracle.nets.sid_forward == the reference's AudioVisualNet (M1/networks.py) on the golden batch: logits, loss,
    selected gradients and the BatchNorm running statistics."""
    from oracle import nets
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in nets.synth_state_dict(nets.sid_shapes(), 3).items()}
    x, lab = torch.tensor(g_nets["x"]), torch.tensor(g_nets["label"])
    stats = {}
    logits = nets.sid_forward(sd, x, lab.shape[1], training=(mode == "train"), stats_out=stats)
    loss = F.binary_cross_entropy_with_logits(logits, lab)
    loss.backward()
    assert float((logits.detach() - torch.tensor(g_nets[f"sid_{mode}_logits"])).abs().max()) < 2e-4
    assert abs(float(loss) - float(g_nets[f"sid_{mode}_loss"])) < 1e-5
    for key in g_nets.files:
        if key.startswith(f"sid_{mode}_grad:"):
            name = key.split(":", 1)[1]
            want = torch.tensor(g_nets[key])
            assert float((sd[name].grad - want).abs().max()) < 2e-3 * float(want.abs().max()) + 1e-7, name


@pytest.mark.parametrize("tag", ["plain", "spread"])
def test_joint_oracle_matches_reference_eval(g_nets, tag):
    """oracle.nets.joint_forward == the reference's JointModel (M2/networks.py), eval mode (train mode runs in the GPU suite,
    where the oracle itself runs on the device; on the CPU it would take minutes)."""
    from oracle import nets, transform as otf
    sd = nets.synth_state_dict(nets.joint_shapes(), 4, spread=(tag == "spread"))
    x, n = torch.tensor(g_nets["x"][:1]), torch.tensor(g_nets["n"][:1])
    with torch.no_grad():
        n_pred, mask = nets.joint_forward(sd, x, n, training=False)
    pre = f"joint_{tag}_eval"
    assert float((n_pred - torch.tensor(g_nets[pre + "_npred"][:1])).abs().max()) < 1e-3 * float(np.abs(g_nets[pre + "_npred"]).max())
    assert float((mask - torch.tensor(g_nets[pre + "_mask"][:1])).abs().mean()) < 1e-5


def test_state_dict_keys_match_reference_counts():
    """The reference's state_dict sizes (84 entries SID, 322 Joint; 2 276 857 + 16 389 372 parameters)."""
    from oracle import nets
    sid, joint = nets.sid_shapes(), nets.joint_shapes()
    assert len(sid) == 84 and len(joint) == 322
    count = lambda sh: sum(int(np.prod(s)) for k, s in sh.items() if "running" not in k and "num_batches" not in k)
    assert count(sid) == 2276857 and count(joint) == 16389372


def test_ssnr_oracle_matches_reference(golden_dir):
    """oracle.metrics vs values produced by the reference's own metrics_ssnr / metrics_ssnr_shift source (M2/metrics.py:86-175)."""
    from oracle import metrics as om, synth
    rows = np.load(os.path.join(golden_dir, "metrics.npz"))["rows"]
    assert len(rows) == 8
    for index, length, eps, ov, seg, ov_s, seg_s in rows:
        c = synth.make_clip(int(index), int(length))
        a = om.metrics_ssnr(c["clean"], c["mixed"], eps=eps)
        b = om.metrics_ssnr_shift(c["clean"], c["mixed"], eps=eps)
        assert abs(a[0] - ov) < 1e-9 and abs(a[1] - seg) < 1e-9 and abs(b[0] - ov_s) < 1e-9 and abs(b[1] - seg_s) < 1e-9


def full_size_inputs(L):
    """The inputs oracle/make_golden_full.py fed to the reference, rebuilt from the same seeds (deterministic numpy)."""
    from oracle import synth, transform as otf
    clips = synth.make_batch(2, length=L, sr=16000 if L == 32000 else 14000, start=40)
    spec = {k: torch.tensor(np.ascontiguousarray(otf.stft_batch(clips[k]))) for k in ("mixed", "noise", "clean", "full_noise")}
    return spec, torch.tensor(clips["label"])


def test_oracle_matches_reference_at_benchmark_size(golden_dir):
    """oracle.nets at the BENCHMARKED clip size (T = 203: 2 s @ 16 kHz, BASELINE configs[1]) against tests/golden/nets_full.npz
    (written by oracle/make_golden_full.py running the reference's own M1 / M2 networks.py, training mode, fp32): logits, losses,
    n_pred, mask and the stored gradient slices -- including the dilation-32 layers, which a T = 71 map never exercises."""
    from oracle import nets, transform as otf
    from oracle.make_golden_full import grad_slice
    g = np.load(golden_dir + "/nets_full.npz")
    spec, lab = full_size_inputs(32000)
    torch.set_num_threads(os.cpu_count() or 1)
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in nets.synth_state_dict(nets.sid_shapes(), 3).items()}
    logits = nets.sid_forward(sd, spec["mixed"], lab.shape[1], training=True)
    loss = F.binary_cross_entropy_with_logits(logits, lab)
    loss.backward()
    assert float((logits.detach() - torch.tensor(g["T203_sid_logits"])).abs().max()) < 2e-4
    assert abs(float(loss.detach()) - float(g["T203_sid_loss"])) < 1e-5
    n = 0
    for key in g.files:
        if key.startswith("T203_sid_grad:"):
            want = torch.tensor(g[key])
            got = torch.tensor(grad_slice(sd[key.split(":", 1)[1]].grad.numpy()))
            assert float((got - want).abs().max()) < 5e-3 * float(want.abs().max()) + 1e-8, key
            n += 1
    assert n >= 9
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in nets.synth_state_dict(nets.joint_shapes(), 4).items()}
    n_pred, mask = nets.joint_forward(sd, spec["mixed"], spec["noise"], training=True)
    rec = otf.batch_fast_icRM_sigmoid(spec["mixed"], mask)
    l1, l2 = F.mse_loss(n_pred, spec["full_noise"]), F.mse_loss(rec, spec["clean"])
    (l1 + l2).backward()
    assert float((mask.detach()[..., ::2] - torch.tensor(g["T203_joint_plain_mask"])).abs().mean()) < 1e-5
    assert float((n_pred.detach()[..., ::2] - torch.tensor(g["T203_joint_plain_npred"])).abs().max()) < 1e-3 * float(np.abs(g["T203_joint_plain_npred"]).max())
    assert abs(float(l1.detach()) - float(g["T203_joint_plain_loss1"])) < 1e-4 and abs(float(l2.detach()) - float(g["T203_joint_plain_loss2"])) < 1e-3
    n = 0
    for key in g.files:
        if key.startswith("T203_joint_plain_grad:"):
            want = torch.tensor(g[key])
            got = torch.tensor(grad_slice(sd[key.split(":", 1)[1]].grad.numpy()))
            assert float((got - want).abs().max()) < 2e-2 * float(want.abs().max()) + 1e-8, key
            n += 1
    assert n >= 18


def test_wss_llr_oracle_matches_reference(golden_dir):
    """oracle.metrics.wss / llr vs the reference's own wss / llr / lpcoeff source (M2/metrics.py:404-681) on three clips."""
    from oracle import metrics as om, synth
    g = np.load(os.path.join(golden_dir, "metrics_lpc.npz"))
    assert len(g.files) == 6
    for key in g.files:
        kind, index, length, srate = key.split(":")
        from oracle.make_golden_metrics import lpc_pair
        ref, deg = lpc_pair(int(index), int(length), int(srate))
        with np.errstate(invalid="ignore"):
            got = om.wss(ref, deg, int(srate)) if kind == "wss" else om.llr(ref, deg, int(srate))
        want = g[key]
        assert got.shape == want.shape and np.array_equal(np.isnan(got), np.isnan(want))
        assert np.nanmax(np.abs(got - want)) < 1e-9 * max(1.0, np.nanmax(np.abs(want)))


def test_resample_oracle_properties(golden_dir):
    """oracle.resample (resampy kaiser_best restated; the real package is absent: this leg is unpinned) against an independent
    polyphase resampler and its own fixture of the reference's demo recording."""
    from scipy.signal import resample_poly
    from oracle import resample as orr
    t = np.arange(4410) / 44100
    x = (0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * np.sin(2 * np.pi * 3000 * t + 1)).astype(np.float32)
    y = orr.resample(x, 44100, 14000)
    z = resample_poly(x.astype(np.float64), 140, 441)
    assert y.shape == (1400,) and y.dtype == np.float32
    assert np.abs(y[200:-200] - z[200:1200]).max() < 2e-3           # (resampy 0.2 truncates its table step: a 0.3 % gain difference)
    g = np.load(os.path.join(golden_dir, "audio_load.npz"))
    w, sr = orr.librosa_load_array(g["stereo_int16"], int(g["sr"]), 14000)
    assert sr == 14000 and np.array_equal(w, g["load_14000"]) and w.shape == (int(np.ceil(5292 * 14000 / 44100)),)
