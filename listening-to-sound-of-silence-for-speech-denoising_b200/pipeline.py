"""In-memory two-stage inference: what the reference runs as three programs joined by JSON + WAV files
(M1/predict.py:116-119 -> M1/create_data_from_pred.py:112,189-221 -> M2/predict.py:303-331,412-426),
as one device-resident call.

    wave (B, L) --STFT--> SID --sigmoid >= 0.5--> bits --gate + STFT--> JointModel --cRM recovery + iSTFT--> denoised (B, 158 (T-1))
"""
import torch

from . import tools, transform


@torch.no_grad()
def denoise(wave, sid, joint, sr=16000, fps=30.0, threshold=0.5, bits=None, frames_per_chunk=None):
    """wave (B, L) CUDA fp32.  sid / joint: the networks of sos_b200.networks in eval mode.  `bits` (B, n) uint8 overrides the
    detector's decision (the reference's `recovered_prediction`); `frames_per_chunk` switches the transforms to their chunked
    long-form variants.  Returns dict(denoised, noise_pred, mask, bits, confidence)."""
    B, L = wave.shape
    ratio = sr / fps
    n_bits = int(L / ratio)                                      # video frames covered by the clip (M1/dataset.py: num_frames)
    stft = (lambda w: transform.stft_chunked(w, frames_per_chunk)) if frames_per_chunk else transform.stft_batch
    mixed = stft(wave)
    conf = None
    if bits is None:
        logits = sid(mixed, n_bits)
        conf = torch.sigmoid(logits)
        bits = tools.logits_to_bits(logits, threshold)           # 1 = non-silent, 0 = silent
    noise_wave = tools.gate_noise(wave, ratio, bits)             # noise_sig = mixed_sig * mask   (M2/predict.py:317)
    noise = stft(noise_wave)
    n_pred, mask = joint(mixed, noise)
    if frames_per_chunk:
        den = transform.istft_chunked(mixed, frames_per_chunk, crm=mask)
    else:
        den = transform.istft_batch(mixed, crm=mask)             # fast_icRM_sigmoid + fast_istft fused (M2/predict.py:422-426)
    return {"denoised": den, "noise_pred": n_pred, "mask": mask, "bits": bits, "confidence": conf}


class GraphedDenoiser(object):
    """`denoise` for a fixed (B, L) captured ONCE into a CUDA graph: the ~500 launches of a forward pass replay as one
    submission, which is what bounds small batches (batch 1: 8.3 ms of launch overhead for ~3 ms of kernels).

        g = GraphedDenoiser(sid, joint, B, L); out = g(wave)      # out: the same dict as denoise(); tensors are reused per call
    """

    def __init__(self, sid, joint, batch, length, sr=16000, fps=30.0, threshold=0.5, frames_per_chunk=None, device=None):
        self.device = device or next(joint.parameters()).device
        self.wave = torch.zeros(batch, length, device=self.device, dtype=torch.float32)
        kw = dict(sr=sr, fps=fps, threshold=threshold, frames_per_chunk=frames_per_chunk)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up off the capture: lazy allocations, attribute settings, tables
            for _ in range(2):
                denoise(self.wave, sid, joint, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = denoise(self.wave, sid, joint, **kw)

    @torch.no_grad()
    def __call__(self, wave):
        self.wave.copy_(wave, non_blocking=True)
        self.graph.replay()
        return self.out


# ----------------------------------------------------------------------------------------------- predict.py-compatible driver
def _read_wav(path):
    """-> (float32 mono waveform in [-1, 1], sample rate) without resampling."""
    import numpy as np
    from scipy.io import wavfile
    sr, x = wavfile.read(path)
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)                                          # librosa.load(mono=True)
    return x, int(sr)


def _write_wav(path, x, sr):
    import numpy as np
    from scipy.io import wavfile
    wavfile.write(path, int(sr), np.asarray(x, dtype=np.float32))   # librosa.output.write_wav writes float32 WAV too


def predict_files(files, sid, joint, out_dir, sr=16000, fps=30.0, threshold=0.5, bit_streams=None, snr=None, save_audio=True):
    """What the reference does with three programs and JSON / WAV hand-offs (M1/predict.py:107-233 -> M1/create_data_from_pred.py:60-270
    -> M2/predict.py:395-530), per file and fully on the device.  Like the reference's prediction phase, the detector sees the WHOLE
    file in one pass (M1/tools.py:328-330 builds ONE item per file with its whole bit stream when pred=True and M1/dataset.py:225-233
    feeds `audio = snd`, so v_num_frames = the number of video frames the file covers, M1/predict.py:117); then gating, denoising,
    and the same artefacts:

      out_dir/eval_results.json           M1/predict.py:185-233: data_total_frames, data_center_frames, sigmoid_threshold, snr,
                                          prediction_statistics {"all": show_metrics}, data[...] (id, path, full_bit_stream, num_frames,
                                          framerate, audio_sample_rate, audio_samples, duration, frame_start_idx, label, pred_label,
                                          match, confidence), sorted by mean confidence, descending
      out_dir/pred_data.json              the hierarchy create_data_from_pred.py writes (dataset_path, num_videos, data_total_frames,
                                          data_center_frames, sigmoid_threshold, snr, prediction_statistics, files[...])
      out_dir/<k>/{noisy_input,noise_intervals,predicted_full_noise,denoised_output}.wav      (M2/predict.py:510-522)

    files: list of WAV paths or (name, waveform ndarray at `sr`) pairs; clips may have different lengths.  WAV files at another
    sample rate (or stereo) go through tools.load_audio = librosa.load(path, sr=sr) (M2/predict.py:303: kaiser_best resampling on the
    device).  bit_streams: optional ground-truth strings ('0' = silent) per file -> prediction_statistics.  The reference's own
    models run at sr = 14000 (DATA_REQUIRED_SR, M1/dataset.py:38); BASELINE.json's configurations at 16000 (the default here).
    Returns the pred_data hierarchy dict."""
    import json
    import os
    from collections import OrderedDict
    import numpy as np
    os.makedirs(out_dir, exist_ok=True)
    dev = next(joint.parameters()).device
    ratio = sr / fps
    groups, labels, preds, stat = [], [], [], []
    for k, f in enumerate(files):
        if isinstance(f, (tuple, list)):
            name, x = f[0], np.asarray(f[1], dtype=np.float32)
            wave = torch.from_numpy(x).to(dev)[None]
        else:
            name = f
            w, _ = tools.load_audio(f, sr=sr, device=dev)           # librosa.load(path, sr=sr)
            wave = w[None]
            x = w.cpu().numpy()
        n_frames = int(len(x) / ratio)
        if n_frames < 1 or 1 + len(x) // 158 < 68:                  # ReflectionPad2d(16) on the T/4 axis (M2/networks.py:176)
            raise RuntimeError(f"{name}: clip of {len(x)} samples is too short for the networks (needs at least 68 STFT frames)")
        out = denoise(wave, sid, joint, sr, fps, threshold)
        bits = out["bits"][0].cpu().numpy()
        conf = out["confidence"][0].cpu().numpy()
        pred = "".join(str(int(b)) for b in bits)
        gt = (bit_streams[k][:n_frames] if bit_streams else None)
        item = OrderedDict([("path", str(name)), ("num_frames", n_frames), ("framerate", fps), ("audio_sample_rate", sr),
                            ("audio_samples", int(len(x))), ("duration", len(x) / sr),
                            ("bit_stream", bit_streams[k] if bit_streams else None), ("predicted_bit_stream", pred),
                            ("recovered_prediction", pred)])
        label = list(gt) if gt else None
        stat.append(OrderedDict([("id", k), ("path", str(name)), ("full_bit_stream", bit_streams[k] if bit_streams else None),
                                 ("num_frames", n_frames), ("framerate", fps), ("audio_sample_rate", sr), ("audio_samples", int(len(x))),
                                 ("duration", len(x) / sr), ("frame_start_idx", 0), ("label", label), ("pred_label", list(pred)),
                                 ("match", label == list(pred) if label else None), ("confidence", [str(c) for c in conf.astype(np.float32)])]))
        if gt:
            labels += [int(c) for c in gt]
            preds += [int(c) for c in pred[:len(gt)]]
        if save_audio:
            d = os.path.join(out_dir, str(k))
            os.makedirs(d, exist_ok=True)
            T = out["mask"].shape[3]
            n = 158 * (T - 1)
            gated = tools.gate_noise(wave, ratio, out["bits"])[0, :n].cpu().numpy()
            pred_noise = transform.istft_batch(out["noise_pred"])[0].cpu().numpy()
            for fname, sig in (("noisy_input", x[:n]), ("noise_intervals", gated), ("predicted_full_noise", pred_noise),
                               ("denoised_output", out["denoised"][0].cpu().numpy())):
                _write_wav(os.path.join(d, fname + ".wav"), sig, sr)
            item["mixed_audio"] = os.path.join(str(k), "noisy_input.wav")
            item["denoised_output"] = os.path.join(str(k), "denoised_output.wav")
        groups.append(item)
    stats = tools.show_metrics(labels, preds) if labels else None
    # ---- eval_results.json (M1/predict.py:185-233)
    stat = sorted(stat, key=lambda it: float(np.mean([float(c) for c in it["confidence"]])), reverse=True)
    eval_results = OrderedDict([("data_total_frames", None), ("data_center_frames", None), ("sigmoid_threshold", threshold), ("snr", snr),
                                ("prediction_statistics", OrderedDict([("all", stats)])), ("data", stat)])
    with open(os.path.join(out_dir, "eval_results.json"), "w") as fp_:
        json.dump(eval_results, fp_, indent=2)
    # ---- pred_data.json (M1/create_data_from_pred.py:212-270)
    paths = [g["path"] for g in groups]
    hierarchy = OrderedDict([("dataset_path", os.path.commonpath(paths) if all(os.path.isabs(p) for p in paths) else ""),
                             ("num_videos", len(groups)), ("data_total_frames", None), ("data_center_frames", None),
                             ("sigmoid_threshold", threshold), ("snr", snr), ("prediction_statistics", stats), ("files", groups)])
    with open(os.path.join(out_dir, "pred_data.json"), "w") as fp_:
        json.dump(hierarchy, fp_, indent=2)
    return hierarchy
