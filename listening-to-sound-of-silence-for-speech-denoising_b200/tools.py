"""Silent-interval bit strings and audio loading on the device (M2/tools.py:340-362, M2/predict.py:232-252,303, M1/predict.py:117-119)."""
import torch

from . import ops


def bits_to_tensor(bit_strings, device):
    """['0101..', ...] ('0' = silent) -> (B, n) uint8."""
    n = len(bit_strings[0])
    assert all(len(b) == n for b in bit_strings), "bit strings of one batch must have equal length"
    return torch.tensor([[1 if c == "1" else 0 for c in b] for b in bit_strings], dtype=torch.uint8, device=device)


def logits_to_bits(logits, threshold=0.5):
    """sigmoid(logit) >= threshold -> 1 (non-silent)  (M1/predict.py:30,117-119)."""
    return (torch.sigmoid(logits) >= threshold).to(torch.uint8)


def convert_bitstreammask_to_audiomask(audio, ratio, bits):
    """Per-sample 0/1 mask (1 = silent) for waveforms (B, L); bits (B, n) uint8 tensor or list of strings."""
    ops.init()
    if not torch.is_tensor(bits):
        bits = bits_to_tensor(bits, audio.device)
    _, mask = ops.gate_wave(audio.contiguous(), bits.contiguous(), ratio, 1, want_mask=True)
    return mask


def gate_noise(mixed, ratio, bits):
    """noise_sig = mixed_sig * mask  (M2/predict.py:317, M2/dataset.py:229)."""
    ops.init()
    if not torch.is_tensor(bits):
        bits = bits_to_tensor(bits, mixed.device)
    return ops.gate_wave(mixed.contiguous(), bits.contiguous(), ratio, 1)


def load_audio(path_or_array, sr=14000, file_sr=None, device=None):
    """librosa.load(path, sr=sr) (M2/predict.py:303, M1/dataset.py:226, M1/create_data_from_pred.py:130): WAV file (or decoded samples
    + file_sr) -> float32 in [-1, 1) -> mono (mean of the channels) -> kaiser_best resampling on the device -> fixed to
    ceil(n * sr / file_sr) samples.  Returns (waveform (n,) CUDA fp32, sr)."""
    import math
    import numpy as np
    ops.init()
    if isinstance(path_or_array, str):
        from scipy.io import wavfile
        file_sr, x = wavfile.read(path_or_array)
    else:
        x = np.asarray(path_or_array)
        assert file_sr is not None, "decoded samples need their sample rate (file_sr)"
    if x.dtype.kind == "i":
        x = x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
    elif x.dtype.kind == "u":
        x = (x.astype(np.float32) - 128.0) / 128.0
    x = x.astype(np.float32)
    if x.ndim == 2:
        x = x.mean(axis=1)                                          # librosa.to_mono
    dev = device or torch.device("cuda", torch.cuda.current_device())
    w = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    if sr is None or int(sr) == int(file_sr):
        return w, int(file_sr)
    n_samples = int(math.ceil(x.shape[0] * float(sr) / file_sr))
    y = ops.resample(w[None], file_sr, sr)[0]
    if y.shape[0] < n_samples:                                      # librosa.util.fix_length
        y = torch.nn.functional.pad(y, (0, n_samples - y.shape[0]))
    return y[:n_samples].contiguous(), int(sr)


def show_metrics(y_true, y_score):
    """M1/tools.py:91-185: frame statistics of a predicted bit stream against the ground truth (1 = non-silent; the SILENT class is
    the positive one), with the reference's key names."""
    from collections import OrderedDict
    import numpy as np
    y_true, y_score = np.int_(y_true), np.int_(y_score)
    count = {int(t): int((y_true == t).sum()) for t in np.unique(y_true)}
    if len(count) == 1:
        count[1 - list(count)[0]] = 0
    base = float((y_true == 1).sum()) / len(y_true)
    accuracy = float((y_true == y_score).sum()) / len(y_true)
    yt, ys = 1 - y_true, 1 - y_score
    tp, fp = int(np.sum(yt * ys)), int(np.sum((yt == 0) * ys))
    tn, fn = int(np.sum((yt == 0) * (ys == 0))), int(np.sum(yt * (ys == 0)))
    div = lambda a, b: (a / b) if b else float("nan")             # (numpy integer division by zero gives nan -> null in the JSON)
    tpr, fpr, precision = div(tp, tp + fn), div(fp, fp + tn), div(tp, tp + fp)
    tnr = 1 - fpr
    f1 = div(2 * tp, 2 * tp + fp + fn)
    auc = (tpr + tnr) / 2
    den = np.sqrt(float(tp + fp) * float(tp + fn) * float(tn + fp) * float(tn + fn))
    mcc = 0 if den == 0 else (tp * tn - fp * fn) / den
    null = lambda v: None if (isinstance(v, float) and v != v) else v
    return OrderedDict([("num_samples", len(y_true)), ("num_silent_samples", count[0]), ("num_non_silent_samples", count[1]), ("base", base),
                        ("accuracy", accuracy), ("true_positive", tp), ("false_positive", fp), ("true_negative", tn), ("false_negative", fn),
                        ("true_pos_rate(recall)", null(float(tpr))), ("false_pos_rate", null(float(fpr))), ("precision", null(float(precision))),
                        ("true_neg_rate", null(float(tnr))), ("f1", null(float(f1))), ("roc_auc", null(float(auc))), ("mcc", null(float(mcc)))])
