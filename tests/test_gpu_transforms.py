"""GPU parity of the HBM-bound kernels (STFT / iSTFT / gating / cRM / losses / Adam / BN / layout) against the CPU oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_stft_matches_oracle(cuda):
    from sos_b200 import transform
    from oracle import synth, transform as otf
    for length in (8000, 28000, 32000):
        clips = synth.make_batch(2, length=length)
        got = transform.stft_batch(torch.tensor(clips["mixed"], device=cuda)).cpu().numpy()
        want = otf.stft_batch(clips["mixed"])
        assert got.shape == want.shape
        err = np.abs(got - want).max()
        assert err < 2e-4, (length, err)          # |S| up to ~50; fp32 dot products of 400 terms


def test_stft_golden_fixture(cuda, golden_dir):
    from sos_b200 import transform
    z = np.load(golden_dir + "/transform.npz")
    got = transform.fast_stft(z["mixed"])
    assert got.shape == z["fast_stft_mixed"].shape
    assert np.abs(got - z["fast_stft_mixed"]).max() < 2e-4
    wav = transform.fast_istft(z["fast_stft_mixed"])
    assert wav.shape == z["fast_istft_mixed"].shape
    assert np.abs(wav - z["fast_istft_mixed"]).max() < 2e-5
    rec = transform.fast_icRM_sigmoid(z["fast_stft_mixed"], z["fast_cRM_sigmoid"])
    ref = z["fast_icRM_sigmoid"]
    assert np.abs(rec - ref).max() < 1e-3 * max(1.0, np.abs(ref).max())


def test_istft_roundtrip_and_fused_crm(cuda):
    from sos_b200 import transform, ops
    from oracle import synth, transform as otf
    clips = synth.make_batch(3, length=32000)
    wave = torch.tensor(clips["mixed"], device=cuda)
    spec = transform.stft_batch(wave)
    back = transform.istft_batch(spec).cpu().numpy()
    assert back.shape == (3, 158 * (spec.shape[3] - 1))
    assert np.abs(back - clips["mixed"][:, :back.shape[1]]).max() < 1e-4          # STFT -> iSTFT round trip
    want = otf.istft_batch(spec.cpu().numpy())
    assert np.abs(back - want).max() < 2e-5
    crm = torch.rand(spec.shape, device=cuda) * 0.5 + 0.25
    fused = transform.istft_batch(spec, crm)
    unfused = transform.istft_batch(ops.icrm_forward(spec, crm))
    assert float((fused - unfused).abs().max()) < 1e-4 * float(unfused.abs().max() + 1)


def test_gating_bit_exact(cuda, golden_dir):
    from sos_b200 import tools
    z = np.load(golden_dir + "/tools.npz")
    n = 0
    for key in z.files:
        if not key.startswith("mask:"):
            continue
        _, length, ratio, bits = key.split(":")
        length, ratio = int(length), float(ratio)
        want = np.unpackbits(z[key])[:length].astype(np.float32)
        audio = torch.zeros(1, length, device=cuda)
        got = tools.convert_bitstreammask_to_audiomask(audio, ratio, [bits]).cpu().numpy()[0]
        assert np.array_equal(got, want), key[:40]
        n += 1
    assert n >= 20


def test_gated_stft_equals_gate_then_stft(cuda):
    from sos_b200 import transform, tools
    from oracle import synth, transform as otf, gating
    clips = synth.make_batch(2, length=32000)
    wave = torch.tensor(clips["mixed"], device=cuda)
    bits = tools.bits_to_tensor(clips["bits"], cuda)
    ratio = 16000 / 30.0
    want = otf.stft_batch(np.stack([gating.gate_noise(w, ratio, b) for w, b in zip(clips["mixed"], clips["bits"])]))
    composed = transform.stft_batch(wave, bits, ratio, 1)                     # gate launch + STFT (the default host path)
    fused = transform.stft_batch(wave, bits, ratio, 1, fused_gate=True)       # mask evaluated inside the STFT kernel (C ABI gate_mode)
    assert np.abs(composed.cpu().numpy() - want).max() < 2e-4
    assert torch.equal(composed, fused)
    clean = transform.stft_batch(wave, bits, ratio, 2, fused_gate=True)       # mode 2: wave * (1 - mask)
    want2 = otf.stft_batch(np.stack([gating.gate_clean(w, ratio, b) for w, b in zip(clips["mixed"], clips["bits"])]))
    assert np.abs(clean.cpu().numpy() - want2).max() < 2e-4


def test_icrm_forward_backward(cuda):
    from sos_b200 import transform
    from oracle import transform as otf
    g = torch.Generator().manual_seed(3)
    Y = torch.randn(2, 2, 256, 19, generator=g)
    crm = torch.rand(2, 2, 256, 19, generator=g) * 0.98 + 0.01
    crm.requires_grad_(True)
    ref = otf.batch_fast_icRM_sigmoid(Y, crm)
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    c2 = crm.detach().to(cuda).requires_grad_(True)
    got = transform.batch_fast_icRM_sigmoid(Y.to(cuda), c2)
    got.backward(go.to(cuda))
    # M = 10*log(crm/(1-crm)) reaches +-46 and |rec| ~ 150: fp32 log/divide differ by a few ulp between libm and the device
    assert float((got.cpu() - ref.detach()).abs().max()) < 2e-5 * float(ref.detach().abs().max())
    assert float((c2.grad.cpu() - crm.grad).abs().max() / crm.grad.abs().max()) < 1e-4


def test_losses_and_adam(cuda):
    from sos_b200 import layers as L, ops
    g = torch.Generator().manual_seed(4)
    a, b = torch.randn(3, 2, 256, 11, generator=g), torch.randn(3, 2, 256, 11, generator=g)
    a.requires_grad_(True)
    ref = torch.nn.MSELoss()(a, b)
    ref.backward()
    ad = a.detach().to(cuda).requires_grad_(True)
    got = L.MSELoss.apply(ad, b.to(cuda))
    got.backward()
    assert abs(float(got) - float(ref)) < 1e-5 * abs(float(ref))
    assert float((ad.grad.cpu() - a.grad).abs().max()) < 1e-7
    x, y = torch.randn(5, 60, generator=g) * 3, (torch.rand(5, 60, generator=g) > 0.5).float()
    x.requires_grad_(True)
    ref = torch.nn.BCEWithLogitsLoss()(x, y)
    ref.backward()
    xd = x.detach().to(cuda).requires_grad_(True)
    got = L.BCEWithLogitsLoss.apply(xd, y.to(cuda))
    got.backward()
    assert abs(float(got) - float(ref)) < 1e-5
    assert float((xd.grad.cpu() - x.grad).abs().max()) < 1e-7
    # Adam: 3 steps against torch.optim.Adam
    p = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) for _ in range(3)]
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    pd, m, v = p.to(cuda), torch.zeros(1000, device=cuda), torch.zeros(1000, device=cuda)
    for i, gr in enumerate(grads):
        pr.grad = gr.clone()
        opt.step()
        ops.adam_step(pd, gr.to(cuda), m, v, 1e-3, i + 1)
    assert float((pd.cpu() - pr.detach()).abs().max()) < 1e-6


def test_adam_device_clock(cuda):
    """sos_adam_step_dev (optimiser clock in device memory, CUDA-graph friendly): 4 steps with a learning-rate change against
    torch.optim.Adam."""
    from sos_b200 import ops
    g = torch.Generator().manual_seed(5)
    p = torch.randn(1000, generator=g)
    grads = [torch.randn(1000, generator=g) * 10 ** (-i) for i in range(4)]
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pr], lr=1e-3)
    pd, m, v = p.to(cuda), torch.zeros(1000, device=cuda), torch.zeros(1000, device=cuda)
    state = torch.tensor([1e-3, 0, 0, 0], device=cuda)
    for i, gr in enumerate(grads):
        if i == 2:
            opt.param_groups[0]["lr"] = 1e-4
            state[0:1].fill_(1e-4)
        pr.grad = gr.clone()
        opt.step()
        ops.adam_step_dev(pd, gr.to(cuda), m, v, state)
    assert float((pd.cpu() - pr.detach()).abs().max()) < 1e-6
    assert float(state[1]) == 4.0


@pytest.mark.parametrize("rows,K,n_out,act", [(203 * 3, 400, 600, 1), (65, 600, 512, 3), (1920, 200, 100, 1), (77, 100, 1, 0), (130, 2048, 96, 0)])
def test_linear_act_matches_fp64(cuda, rows, K, n_out, act):
    """layers.LinearAct (nn.Linear + ReLU / Sigmoid of the heads, M1/networks.py:96-98, M2/networks.py:65-70) = one 3-tap TF32 tap-GEMM
    launch over split operands: fp32-grade.  Forward, input gradient, weight and bias gradients against float64."""
    from sos_b200 import layers as L
    g = torch.Generator().manual_seed(rows + K)
    x = torch.randn(rows, K, generator=g, dtype=torch.float64)
    W = torch.randn(n_out, K, generator=g, dtype=torch.float64) / K ** 0.5
    b = torch.randn(n_out, generator=g, dtype=torch.float64) * 0.1
    go = torch.randn(rows, n_out, generator=g, dtype=torch.float64)
    xr, Wr, br = (t.clone().requires_grad_(True) for t in (x, W, b))
    pre = xr @ Wr.t() + br
    y = torch.relu(pre) if act == 1 else (torch.sigmoid(pre) if act == 3 else pre)
    y.backward(go)
    rel = lambda a, r: float((a.double().cpu() - r).abs().max() / (r.abs().max() + 1e-30))
    tol = 2e-5 * max(1.0, K / 400)           # (the tensor core truncates its fp32 accumulation: ~2^-24 per K step; plain TF32 would be ~1e-3)
    # default mode: fp32-grade forward, ONE TF32 pass for the data / weight gradients (layers._bw_passes; operands rounded to 11 bits:
    # ~1e-3 of scale); precise mode: three passes everywhere
    for precise in (False, True):
        old = L.set_precise(precise)
        try:
            xd, Wd, bd = (t.float().to(cuda).requires_grad_(True) for t in (x, W, b))
            got = L.LinearAct.apply(xd.view(rows, 1, K), Wd, bd, act)
            assert got.shape == (rows, 1, n_out)
            got.backward(go.float().to(cuda).view(rows, 1, n_out))
        finally:
            L.set_precise(old)
        errs = (rel(got.detach().view(rows, n_out), y.detach()), rel(xd.grad, xr.grad), rel(Wd.grad, Wr.grad), rel(bd.grad, br.grad))
        print("linear_act", "precise" if precise else "default", rows, K, n_out, act, ["%.2e" % e for e in errs])
        gtol = 2.5 * tol if precise else 1.5e-3
        assert errs[0] < tol and errs[1] < gtol and errs[2] < gtol * (max(1.0, rows / K) if precise else 1.0) and errs[3] < 2.5 * tol, errs


def test_seq_map_roundtrip(cuda):
    from sos_b200 import layers as L
    h = torch.randn(37, 5, 70, device=cuda, requires_grad=True)
    m = L.SeqToMap.apply(h)
    assert torch.equal(m.detach(), h.detach().permute(1, 2, 0).contiguous())
    go = torch.randn_like(m)
    m.backward(go)
    assert torch.equal(h.grad, go.permute(2, 0, 1).contiguous())


@pytest.mark.parametrize("act", [1, 2])
@pytest.mark.parametrize("C", [48, 96, 8, 256])
def test_bn_act_train(cuda, act, C):
    from sos_b200 import layers as L, ops
    g = torch.Generator().manual_seed(C + act)
    y = torch.randn(2, C, 12, 17, generator=g) * 2 + 0.5
    bn = torch.nn.BatchNorm2d(C)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.1)
    prelu = torch.nn.PReLU()
    y.requires_grad_(True)
    ref = bn(y)
    ref = torch.relu(ref) if act == 1 else prelu(ref)
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    bn2 = torch.nn.BatchNorm2d(C).to(cuda)
    with torch.no_grad():
        bn2.weight.copy_(bn.weight.detach())
        bn2.bias.copy_(bn.bias.detach())
    slope = torch.nn.Parameter(torch.tensor([0.25], device=cuda)) if act == 2 else None
    yd = ops.nchw_to_nhwc(y.detach().to(cuda), C).requires_grad_(True)
    z = L.bn_act(yd, bn2, act, slope, True, round_out=False, round_grad=False)
    z.backward(ops.nchw_to_nhwc(go.to(cuda), C))
    assert float((ops.nhwc_to_nchw(z, C).cpu() - ref.detach()).abs().max()) < 2e-5
    assert float((ops.nhwc_to_nchw(yd.grad, C).cpu() - y.grad).abs().max()) < 1e-4 * float(y.grad.abs().max() + 1)
    assert float((bn2.weight.grad.cpu() - bn.weight.grad).abs().max()) < 1e-3 * float(bn.weight.grad.abs().max() + 1)
    assert float((bn2.bias.grad.cpu() - bn.bias.grad).abs().max()) < 1e-3 * float(bn.bias.grad.abs().max() + 1)
    assert float((bn2.running_mean.cpu() - bn.running_mean).abs().max()) < 1e-6
    assert float((bn2.running_var.cpu() - bn.running_var).abs().max()) < 1e-5
    if act == 2:
        assert abs(float(slope.grad) - float(prelu.weight.grad)) < 1e-3 * (abs(float(prelu.weight.grad)) + 1)


def test_padcat_reflect_resize(cuda):
    from sos_b200 import layers as L, ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(8)
    a = torch.randn(2, 64, 16, 22, generator=g, requires_grad=True)       # will be resized to 16 x 21
    b = torch.randn(2, 64, 16, 21, generator=g, requires_grad=True)
    ref = F.pad(torch.cat([F.interpolate(a, (16, 21)), b], 1), (3, 3, 3, 3), mode="reflect")
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    ad = ops.nchw_to_nhwc(a.detach().to(cuda), 64).requires_grad_(True)
    bd = ops.nchw_to_nhwc(b.detach().to(cuda), 64).requires_grad_(True)
    out = L.PadCat.apply(3, 16, 21, ad, bd)
    out.backward(ops.nchw_to_nhwc(go.to(cuda), 128))
    assert float((ops.nhwc_to_nchw(out, 128).cpu() - ref.detach()).abs().max()) == 0.0
    assert float((ops.nhwc_to_nchw(ad.grad, 64).cpu() - a.grad).abs().max()) < 1e-5
    assert float((ops.nhwc_to_nchw(bd.grad, 64).cpu() - b.grad).abs().max()) < 1e-5


@pytest.mark.parametrize("pad,H,W", [(1, 16, 21), (2, 9, 7), (3, 16, 21), (2, 3, 5), (1, 2, 2)])
def test_padcat_reflect_same_size(cuda, pad, H, W):
    """Same-size sources: backward = sos_copy_view_fold (fold + un-pad + channel slice in one pass over the untouched padded map)."""
    from sos_b200 import layers as L, ops
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(9)
    a = torch.randn(2, 8, H, W, generator=g, requires_grad=True)
    b = torch.randn(2, 16, H, W, generator=g, requires_grad=True)
    ref = F.pad(torch.cat([a, b], 1), (pad, pad, pad, pad), mode="reflect")
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    ad = ops.nchw_to_nhwc(a.detach().to(cuda), 8).requires_grad_(True)
    bd = ops.nchw_to_nhwc(b.detach().to(cuda), 16).requires_grad_(True)
    out = L.PadCat.apply(pad, H, W, ad, bd)
    gd = ops.nchw_to_nhwc(go.to(cuda), 24)
    keep = gd.clone()
    out.backward(gd)
    assert torch.equal(gd, keep)                                           # the incoming gradient map is not modified
    assert float((ops.nhwc_to_nchw(out, 24).cpu() - ref.detach()).abs().max()) == 0.0
    assert float((ops.nhwc_to_nchw(ad.grad, 8).cpu() - a.grad).abs().max()) < 1e-5
    assert float((ops.nhwc_to_nchw(bd.grad, 16).cpu() - b.grad).abs().max()) < 1e-5


@pytest.mark.parametrize("I,H,T,B", [(96, 40, 13, 5), (64, 200, 37, 40), (32, 100, 60, 3), (48, 200, 29, 32), (16, 100, 17, 1), (24, 250, 9, 7)])
def test_bilstm_matches_torch(cuda, I, H, T, B):
    """nn.LSTM(bidirectional) forward / backward.  Batches of at most 32 clips run the cluster kernels (one thread-block cluster per
    direction, h / dh exchanged through distributed shared memory: 3, 7, 13 and 16 CTAs here); B = 40 runs the global-barrier
    kernels over two batch tiles."""
    from sos_b200 import networks
    torch.manual_seed(2)
    ref = torch.nn.LSTM(I, H, bidirectional=True)
    mine = networks.BiLSTM(I, H, bidirectional=True)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(cuda)
    x = torch.randn(T, B, I, requires_grad=True)
    out, _ = ref(x)
    go = torch.randn(out.shape)
    out.backward(go)
    from sos_b200 import layers as L
    # (the projections are split-TF32 tensor-core GEMMs: exact products, but the tensor core's fp32 accumulation truncates instead of
    #  rounding, ~2^-24 per K step: 1e-5 relative at K = 1600.  The data / weight gradients take ONE TF32 pass in the default mode
    #  (layers._bw_passes: ~1e-3 of scale) and three in precise mode)
    for precise in (False, True):
        old = L.set_precise(precise)
        try:
            for q in mine.parameters():
                q.grad = None
            xd = x.detach().to(cuda).requires_grad_(True)
            got = mine(xd)
            got.backward(go.to(cuda))
        finally:
            L.set_precise(old)
        assert float((got.cpu() - out.detach()).abs().max()) < 2e-5
        gx, gw = (5e-5, 1e-4) if precise else (2e-3, 2e-3)
        assert float((xd.grad.cpu() - x.grad).abs().max()) < gx * float(x.grad.abs().max() + 1)
        for (k, p), (_, q) in zip(ref.named_parameters(), mine.named_parameters()):
            assert float((q.grad.cpu() - p.grad).abs().max()) < gw * float(p.grad.abs().max() + 1), (k, precise)


def test_longform_chunked_equals_unchunked(cuda):
    """BASELINE configs[4]: 10 s clips (L = 160000, T = 1013): the chunked STFT and the streaming overlap-add are bit-identical
    to the one-shot transforms, also for ragged chunk sizes and with the fused cRM recovery."""
    from sos_b200 import transform
    g = torch.Generator().manual_seed(21)
    wave = (torch.randn(2, 160000, generator=g) * 0.2).to(cuda)
    whole = transform.stft_batch(wave)
    assert whole.shape == (2, 2, 256, 1013)
    for fpc in (256, 300, 1013, 97):
        assert torch.equal(transform.stft_chunked(wave, fpc), whole), fpc
    w_whole = transform.istft_batch(whole)
    crm = (torch.rand(whole.shape, generator=g) * 0.9 + 0.05).to(cuda)
    w_crm = transform.istft_batch(whole, crm)
    for fpc in (256, 300, 1012, 97):
        assert torch.equal(transform.istft_chunked(whole, fpc), w_whole), fpc
        assert torch.equal(transform.istft_chunked(whole, fpc, crm=crm), w_crm), fpc
    # and against the oracle at this length
    from oracle import transform as otf
    ref = otf.stft_batch(wave[:1].cpu().numpy())
    assert float((whole[:1].cpu() - torch.tensor(ref)).abs().max()) < 3e-4


def test_denoise_pipeline_matches_oracle(cuda):
    """The in-memory inference pipeline (STFT -> SID -> threshold -> gate -> STFT -> JointModel -> cRM + iSTFT) against the CPU
    oracle run stage by stage.  The detector's 0.5 threshold is a discontinuity, so the oracle's downstream stages are fed the
    bits the device produced; bits are compared wherever the oracle's confidence is not within 0.02 of the threshold."""
    from sos_b200 import networks, pipeline
    from oracle import gating, nets, synth, transform as otf
    L, sr, fps = 12640, 16000, 30.0
    clips = synth.make_batch(2, length=L)
    wave = torch.tensor(clips["mixed"], device=cuda)
    sid = networks.get_network()
    sid.load_state_dict(nets.synth_state_dict(nets.sid_shapes(), 3))
    joint = networks.get_network(object())
    joint.load_state_dict(nets.synth_state_dict(nets.joint_shapes(), 4))
    sid, joint = sid.to(cuda).eval(), joint.to(cuda).eval()
    out = pipeline.denoise(wave, sid, joint, sr, fps)
    n_bits = int(L / (sr / fps))
    assert out["bits"].shape == (2, n_bits) and out["denoised"].shape == (2, 158 * (L // 158))
    # oracle, stage by stage
    mixed = torch.tensor(otf.stft_batch(clips["mixed"]))
    with torch.no_grad():
        logits = nets.sid_forward(nets.synth_state_dict(nets.sid_shapes(), 3), mixed, n_bits, training=False)
    conf = torch.sigmoid(logits)
    assert float((out["confidence"].cpu() - conf).abs().max()) < 5e-3
    sure = (conf - 0.5).abs() > 0.02
    assert bool(((out["bits"].cpu() > 0) == (conf >= 0.5))[sure].all())
    bit_strings = ["".join(str(int(b)) for b in row) for row in out["bits"].cpu().numpy()]
    noise_w = np.stack([gating.gate_noise(clips["mixed"][i], sr / fps, bit_strings[i]) for i in range(2)]).astype(np.float32)
    noise = torch.tensor(otf.stft_batch(noise_w))
    with torch.no_grad():
        n_pred, mask = nets.joint_forward(nets.synth_state_dict(nets.joint_shapes(), 4), mixed, noise, training=False)
    rec = otf.batch_fast_icRM_sigmoid(mixed, mask)
    den = otf.istft_batch(rec.numpy())
    e_mask = float((out["mask"].cpu() - mask).abs().mean())
    e_wave = float(np.abs(out["denoised"].cpu().numpy() - den).mean())
    print(f"denoise pipeline: mask L1 {e_mask:.2e}, waveform L1 {e_wave:.2e} (|wave| mean {np.abs(den).mean():.3f})")
    assert e_mask < 1e-3                                            # north_star bar
    assert e_wave < 2e-2 * float(np.abs(den).mean()) + 1e-4         # cRM recovery amplifies mask error x40 at crm = 0.5


def test_graphed_denoiser_matches_eager(cuda):
    """The CUDA-graph replay of the inference pipeline returns what the eager call returns, for changing inputs."""
    from sos_b200 import networks, pipeline
    from oracle import synth
    torch.manual_seed(0)
    sid = networks.get_network().to(cuda).eval()
    torch.manual_seed(1)
    joint = networks.get_network(object()).to(cuda).eval()
    L = 16000
    g = pipeline.GraphedDenoiser(sid, joint, 2, L, 16000, 30.0)
    for seed in (0, 5):
        wave = torch.tensor(synth.make_batch(2, length=L, start=seed)["mixed"], device=cuda)
        ref = pipeline.denoise(wave, sid, joint, 16000, 30.0)
        out = g(wave)
        torch.cuda.synchronize()
        assert torch.equal(out["bits"], ref["bits"])
        assert torch.equal(out["denoised"], ref["denoised"]) and torch.equal(out["mask"], ref["mask"])


def test_device_item_builder_matches_reference_recipe(cuda):
    """the stage-2 training item built on the device (gate, add_signals at an SNR, gate, 4 x STFT, cRM target) against
    the oracle's restatement of M2/dataset.py:167-259 / M2/tools.py:217-276, incl. a silent clip and norm = 0."""
    from sos_b200 import datapipe, ops
    from oracle import synth, transform as otf
    from oracle.gating import bits_to_sample_mask
    rng = np.random.default_rng(7)
    B, L, sr, fps = 4, 16000, 16000, 30.0
    n_bits = int(L / (sr / fps))
    audio = (rng.standard_normal((B, L)) * 0.1).astype(np.float32)
    noise = (rng.standard_normal((B, L)) * 0.3).astype(np.float32)
    audio[2] = 0                                                   # signal_power == 0: the noise is added as it is
    bits = ["".join(rng.choice(["0", "1"], p=[0.4, 0.6]) for _ in range(n_bits)) for _ in range(B)]
    snrs = [-7.0, 3.0, 0.0, 10.0]
    for norm in (0.5, 0.0):
        item = datapipe.make_joint_items(torch.tensor(audio, device=cuda), torch.tensor(noise, device=cuda), snrs, bits, sr, fps, norm=norm,
                                         want_waves=True)
        torch.cuda.synchronize()
        for i in range(B):
            mask = bits_to_sample_mask(L, sr / fps, bits[i]).astype(np.float32)
            mixed, clean, full = synth.add_signals(audio[i] * (1 - mask), noise[i], snrs[i], norm=norm)
            gated = mixed * mask
            for name, ref in (("mixed", mixed), ("clean", clean), ("full_noise", full), ("noise", gated)):
                got = item["waves"][name][i].cpu().numpy()
                assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max()), (norm, i, name, np.abs(got - ref).max())
                spec = item[name][i].cpu().numpy()
                assert np.abs(spec - otf.fast_stft(ref).transpose(2, 0, 1)).max() < 3e-4, (norm, i, name)
            crm = otf.fast_cRM_sigmoid(otf.fast_stft(clean), otf.fast_stft(mixed)).transpose(2, 0, 1)
            # the target divides by |Y|^2 + 1e-8: compare where the mixture bin is not vanishing (elsewhere the ratio amplifies
            # the 1e-4 transform tolerance without bound)
            Y = otf.fast_stft(mixed)
            ok = (Y[..., 0] ** 2 + Y[..., 1] ** 2) > 1e-2
            err = np.abs(item["mask"][i].cpu().numpy() - crm)[:, ok]
            assert err.max() < 2e-3, (norm, i, err.max())
    sid_item = datapipe.make_sid_items(torch.tensor(audio, device=cuda), torch.tensor(noise, device=cuda), snrs, bits, sr, fps)
    assert sid_item["label"].shape == (B, n_bits) and sid_item["audio"].shape == (B, 2, 256, 1 + L // 158)
    assert torch.equal(sid_item["audio"], datapipe.make_joint_items(torch.tensor(audio, device=cuda), torch.tensor(noise, device=cuda), snrs,
                                                                   bits, sr, fps)["mixed"])


def test_predict_files_driver(cuda, tmp_path):
    """the predict.py-shaped driver on two clips of different lengths: pred_data.json keys, bit strings of the right
    length, and denoised_output.wav equal to the in-memory pipeline's waveform."""
    import json
    from scipy.io import wavfile
    from sos_b200 import networks, pipeline
    from oracle import synth
    torch.manual_seed(0)
    sid = networks.get_network().to(cuda).eval()
    torch.manual_seed(1)
    joint = networks.get_network(object()).to(cuda).eval()
    clips = [synth.make_clip(0, 32000), synth.make_clip(1, 48000)]
    paths = []
    for i, c in enumerate(clips):
        p = str(tmp_path / f"clip{i}.wav")
        wavfile.write(p, 16000, c["mixed"].astype(np.float32))
        paths.append(p)
    h = pipeline.predict_files(paths, sid, joint, str(tmp_path / "out"), 16000, 30.0, bit_streams=[c["bits"] for c in clips])
    saved = json.load(open(tmp_path / "out" / "pred_data.json"))
    assert list(saved) == ["dataset_path", "num_videos", "data_total_frames", "data_center_frames", "sigmoid_threshold", "snr",
                           "prediction_statistics", "files"]
    assert saved["num_videos"] == 2 and list(saved["prediction_statistics"])[:5] == ["num_samples", "num_silent_samples", "num_non_silent_samples", "base", "accuracy"]
    assert saved["prediction_statistics"]["num_samples"] == sum(int(len(c["mixed"]) / (16000 / 30.0)) for c in clips) and "mcc" in saved["prediction_statistics"]
    ev = json.load(open(tmp_path / "out" / "eval_results.json"))                  # M1/predict.py:185-233
    assert list(ev) == ["data_total_frames", "data_center_frames", "sigmoid_threshold", "snr", "prediction_statistics", "data"]
    assert list(ev["prediction_statistics"]) == ["all"] and ev["prediction_statistics"]["all"] == saved["prediction_statistics"]
    assert [list(d) for d in ev["data"]] == [["id", "path", "full_bit_stream", "num_frames", "framerate", "audio_sample_rate", "audio_samples",
                                              "duration", "frame_start_idx", "label", "pred_label", "match", "confidence"]] * 2
    means = [np.mean([float(c) for c in d["confidence"]]) for d in ev["data"]]
    assert means[0] >= means[1] and all(len(d["confidence"]) == d["num_frames"] == len(d["pred_label"]) for d in ev["data"])
    assert all(0.0 <= float(c) <= 1.0 for d in ev["data"] for c in d["confidence"])
    for i, (c, f) in enumerate(zip(clips, saved["files"])):
        n = int(len(c["mixed"]) / (16000 / 30.0))
        assert f["num_frames"] == n and len(f["predicted_bit_stream"]) == n and f["recovered_prediction"] == f["predicted_bit_stream"]
        assert set(f["predicted_bit_stream"]) <= {"0", "1"}
        sr, den = wavfile.read(tmp_path / "out" / str(i) / "denoised_output.wav")
        ref = pipeline.denoise(torch.tensor(c["mixed"], device=cuda)[None], sid, joint, 16000, 30.0)["denoised"][0].cpu().numpy()
        assert sr == 16000 and den.shape == ref.shape and np.array_equal(den, ref)
    assert h["files"][0]["path"] == paths[0]


def test_to_half_saturates_and_scales(cuda):
    """Half operand producers: round to nearest even, zero-padded channels, saturation at +-65504 (no infinities inside a GEMM
    operand), and the power-of-two scale of gradient-sized values (exact, RMS brought to ~1)."""
    from sos_b200 import ops
    x = torch.tensor([[1e6, -1e6, 65504.0, 1.0, 0.1, -3.0e-5, 0.0, 2049.0]], device=cuda)
    h = ops.to_half(x, 16)
    assert h.shape == (1, 16) and h.dtype == torch.float16
    want = torch.tensor([65504.0, -65504.0, 65504.0, 1.0, 0.1, -3.0e-5, 0.0, 2049.0]).half()      # 2049 -> 2048 (ties to even)
    assert torch.equal(h[0, :8].cpu(), want) and float(h[0, 8:].abs().max()) == 0
    g = torch.randn(4096, 8, device=cuda) * 3e-9
    gh, scal = ops.to_half(g, scaled=True)
    s, inv = float(scal[0]), float(scal[1])
    assert s * inv == 1.0 and abs(np.log2(s) - round(np.log2(s))) < 1e-12                           # an exact power of two
    rms = float((gh.float() ** 2).mean().sqrt())
    assert 0.5 < rms < 2.0
    back = gh.float() * inv
    assert float((back - g).abs().max()) <= 2.0 ** -11 * float(g.abs().max())                       # 11-bit significand kept


def test_ssnr_metrics_match_reference(cuda, golden_dir):
    """segmental / overall SNR on the device vs the golden values from the reference's own source and vs the oracle,
    single waveforms and a batch, clips of three lengths."""
    import os
    from sos_b200 import metrics
    from oracle import metrics as om, synth
    rows = np.load(os.path.join(golden_dir, "metrics.npz"))["rows"]
    for index, length, eps, ov, seg, ov_s, seg_s in rows:
        c = synth.make_clip(int(index), int(length))
        a = metrics.metrics_ssnr(c["clean"], c["mixed"], eps=eps)
        b = metrics.metrics_ssnr_shift(c["clean"], c["mixed"], eps=eps)
        assert abs(a[0] - ov) < 2e-4 and abs(a[1] - seg) < 2e-4 and abs(b[0] - ov_s) < 2e-4 and abs(b[1] - seg_s) < 2e-4, (index, a, b)
    batch = synth.make_batch(3, length=24000)
    ov, seg = metrics.metrics_ssnr(torch.tensor(batch["clean"], device=cuda), torch.tensor(batch["mixed"], device=cuda))
    for i in range(3):
        want = om.metrics_ssnr(batch["clean"][i], batch["mixed"][i])
        assert abs(float(ov[i]) - want[0]) < 2e-4 and abs(float(seg[i]) - want[1]) < 2e-4
    l1 = metrics.metrics_L1(batch["mixed"][0], batch["clean"][0])
    assert abs(l1 - float(np.mean(np.abs(batch["mixed"][0] - batch["clean"][0])))) < 1e-7


def test_waveform_dataset_collate(cuda):
    """WaveformDataset.collate: raw crops -> the reference's item dict on the device; the joint agent trains on it unchanged."""
    from sos_b200 import agent as ag, datapipe
    from oracle import synth
    clips = []
    for i in range(2):
        c = synth.make_clip(i, 16000)
        clips.append({"audio": c["clean"], "bitstream": c["bits"]})
    noises = [np.random.default_rng(1).standard_normal(40000).astype(np.float32) * 0.2]
    ds = datapipe.WaveformDataset(clips, noises, seed=0)
    item = ds.collate([ds[0], ds[1]])
    assert set(item) == {"mixed", "clean", "noise", "full_noise", "mask", "start", "bitstream"}
    assert item["mixed"].shape == (2, 2, 256, 1 + 16000 // 158) and item["mask"].shape == item["mixed"].shape
    assert float(item["mask"].min()) > 0 and float(item["mask"].max()) < 1
    torch.manual_seed(1)
    joint = ag.MyAgent(ag.default_config(model="joint", sr=16000, fps=30.0))
    _, losses = joint.train_func(item)
    assert all(torch.isfinite(v) for v in losses.values())
    sid_item = ds.collate([ds[0], ds[1]], model="sid")
    assert sid_item["label"].shape == (2, 30) and sid_item["audio"].shape == item["mixed"].shape


@pytest.mark.parametrize("sr_in,sr_out,n", [(44100, 14000, 9000), (44100, 16000, 9000), (8000, 16000, 3000), (22050, 14000, 5000)])
def test_resample_matches_oracle(cuda, sr_in, sr_out, n):
    """ops.resample (librosa.load's kaiser_best resampler, M2/predict.py:303) against the numpy restatement of resampy's published
    algorithm, tap for tap (the same float32 accumulator rounding): <= 2e-7 absolute on signals of amplitude 0.5."""
    from sos_b200 import ops
    from oracle import resample as orr
    rng = np.random.default_rng(sr_in + sr_out)
    t = np.arange(n) / sr_in
    x = np.stack([0.3 * np.sin(2 * np.pi * 440 * t) + 0.2 * rng.standard_normal(n), 0.5 * rng.standard_normal(n) * (rng.random(n) > 0.3)]).astype(np.float32)
    got = ops.resample(torch.tensor(x, device=cuda), sr_in, sr_out).cpu().numpy()
    for b in range(2):
        want = orr.resample(x[b], sr_in, sr_out)
        assert got[b].shape == want.shape == (int(n * sr_out / sr_in),)
        assert np.abs(got[b] - want).max() < 2e-7, np.abs(got[b] - want).max()


def test_load_audio_on_reference_recording(cuda, golden_dir, tmp_path):
    """tools.load_audio = librosa.load(path, sr=14000 / 16000) on 0.12 s of the reference's own demo recording
    (data/sounds_of_silence_audioonly_original/sos_1.wav: 44.1 kHz, stereo, int16; fixture tests/golden/audio_load.npz with the
    oracle's output): int16 -> float, mono mix, resampling, fix_length."""
    from scipy.io import wavfile
    from sos_b200 import tools
    g = np.load(golden_dir + "/audio_load.npz")
    path = str(tmp_path / "seg.wav")
    wavfile.write(path, int(g["sr"]), g["stereo_int16"])
    for tgt in (14000, 16000):
        y, sr = tools.load_audio(path, sr=tgt)
        want = g[f"load_{tgt}"]
        assert sr == tgt and tuple(y.shape) == want.shape
        assert np.abs(y.cpu().numpy() - want).max() < 2e-7
    y, sr = tools.load_audio(path, sr=None)
    assert sr == 44100 and y.shape[0] == g["stereo_int16"].shape[0]


def test_wss_llr_match_reference(cuda, golden_dir):
    """metrics.wss / metrics.llr (M2/metrics.py:404-681) against values produced by the reference's own source
    (tests/golden/metrics_lpc.npz, oracle/make_golden_metrics.py) at 16 / 14 / 8 kHz.  WSS is double arithmetic on both sides:
    2e-4 relative with float32 inputs.  LLR: the reference casts the autocorrelation and the LPC vector to float32 before its two
    quadratic forms; the kernel applies the same casts: <= 1e-4 absolute (observed 4e-7).  (On perfectly predictable signals -- the
    noise-free synthetic harmonics -- those float32 forms cancel to rounding noise in the reference itself; the fixtures add noise.)"""
    from sos_b200 import metrics
    from oracle import synth
    g = np.load(golden_dir + "/metrics_lpc.npz")
    n_llr = 0
    for key in g.files:
        kind, index, length, srate = key.split(":")
        index, length, srate = int(index), int(length), int(srate)
        from oracle.make_golden_metrics import lpc_pair
        ref, deg = (w.astype(np.float32) for w in lpc_pair(index, length, srate))
        want = g[key]
        if kind == "wss":
            got = metrics.wss(ref, deg, srate)
            assert got.shape == want.shape
            assert np.abs(got - want).max() < 2e-4 * np.abs(want).max(), (key, np.abs(got - want).max())     # (inputs rounded to float32 here)
        else:
            got = metrics.llr(ref, deg, srate)
            assert got.shape == want.shape
            ok = np.isfinite(want) & np.isfinite(got)
            err = np.abs(got[ok] - want[ok])
            print(key, "frames", len(want), "finite in both", int(ok.sum()), "nan only here", int((np.isfinite(want) & ~np.isfinite(got)).sum()),
                  "nan only there", int((~np.isfinite(want) & np.isfinite(got)).sum()), "median err", float(np.median(err)), "max err", float(err.max()))
            assert ok.sum() == len(want), (key, ok.sum())
            assert err.max() < 1e-4, (key, err.max())
            n_llr += int(ok.sum())
    assert n_llr > 100
    both = metrics.wss(np.stack([ref, ref]), np.stack([deg, ref]), srate)        # batched: second pair is identical -> distance 0
    assert tuple(both.shape) == (2, len(want)) and float(both[1].abs().max()) < 1e-9
    sig, bak, ovl, pesq, seg, snr = metrics.CompositeEval(ref, deg, srate, pesq_raw=2.5)
    assert 1 <= sig <= 5 and 1 <= bak <= 5 and 1 <= ovl <= 5 and pesq == 2.5 and np.isfinite(seg) and np.isfinite(snr)
