"""Training items built ON THE DEVICE from waveform crops.

The reference builds every item in CPU DataLoader workers (60-70 processes): bit string -> sample mask, clean = audio * (1 - mask),
add_signals at a random SNR with norm 0.5, noise_sig = mixed * mask, four librosa STFTs and the cRM target
(M2/dataset.py:144-320; M1/dataset.py:215-352).  Here a Dataset only has to deliver the raw crops (clean speech, a noise
interval, the ground-truth bits, an SNR); everything else is six kernel launches for the whole batch, and the item dicts keep the
reference's keys and layouts so the agents' `forward(data)` takes them unchanged.
"""
import random

import torch

from . import ops, tools, transform

SNRS = [-10, -7, -3, 0, 3, 7, 10]          # M2/dataset.py:34


def crop_noise(noises, length, batch, device, rng=random):
    """add_noise_to_audio's random interval of a random noise track (M2/tools.py:279-292, M2/dataset.py:206) for every clip of
    the batch.  noises: list of 1-D CUDA/CPU tensors, each at least `length` long."""
    out = torch.empty(batch, length, device=device, dtype=torch.float32)
    for i in range(batch):
        n = rng.choice(noises)
        if n.numel() < length:
            raise ValueError(f"noise track of {n.numel()} samples is shorter than the {length}-sample clip")
        start = rng.randint(0, n.numel() - length)
        out[i].copy_(n[start:start + length], non_blocking=True)
    return out


def make_joint_items(audio, noise, snr_db, bits, sr=16000, fps=30.0, norm=0.5, want_waves=False):
    """The stage-2 item dict (M2/dataset.py:311-320) for a whole batch.
    audio, noise (B, L) CUDA fp32 crops; snr_db (B,) floats (tensor or list); bits (B, n) uint8, 0 = silent (or bit strings)."""
    ops.init()
    dev = audio.device
    if not torch.is_tensor(bits):
        bits = tools.bits_to_tensor(bits, dev)
    snr = torch.as_tensor(snr_db, dtype=torch.float32, device=dev)
    ratio = sr / fps
    B = audio.shape[0]
    audio = ops.gate_wave(audio.contiguous(), bits.contiguous(), ratio, 2)       # audio * (1 - mask)        M2/dataset.py:193
    mixed_w, clean_w, full_w = ops.add_signals(audio, noise.contiguous(), snr, norm)      # M2/dataset.py:217
    noise_w = ops.gate_wave(mixed_w, bits, ratio, 1)                               # noise_sig = mixed * mask   M2/dataset.py:229
    spec = transform.stft_batch(torch.cat([mixed_w, clean_w, noise_w, full_w]))    # M2/dataset.py:234-237, one launch
    mixed, clean, noise_s, full = spec[:B], spec[B:2 * B], spec[2 * B:3 * B], spec[3 * B:]
    item = {"mixed": mixed, "clean": clean, "noise": noise_s, "full_noise": full,
            "mask": ops.crm_forward(clean, mixed),                                 # fast_cRM_sigmoid           M2/dataset.py:239
            "start": 0, "bitstream": bits}
    if want_waves:
        item["waves"] = {"mixed": mixed_w, "clean": clean_w, "noise": noise_w, "full_noise": full_w}
    return item


def make_sid_items(audio, noise, snr_db, bits, sr=16000, fps=30.0, norm=0.5, clean_audio=True):
    """The stage-1 item dict (M1/dataset.py:348-352): {"label": (B, n) float 1 = non-silent, "audio": (B, 2, 256, T)}."""
    ops.init()
    dev = audio.device
    if not torch.is_tensor(bits):
        bits = tools.bits_to_tensor(bits, dev)
    ratio = sr / fps
    if clean_audio:                                                                # M1/dataset.py:243-262
        audio = ops.gate_wave(audio.contiguous(), bits.contiguous(), ratio, 2)
        audio, _, _ = ops.add_signals(audio, noise.contiguous(), torch.as_tensor(snr_db, dtype=torch.float32, device=dev), norm)
    return {"label": bits.float(), "audio": transform.stft_batch(audio)}


class WaveformDataset(torch.utils.data.Dataset):
    """The waveform-only Dataset of SURVEY.md 8f-2: items are RAW crops (what the reference's `__getitem__` has after
    `librosa.load` and slicing, M2/dataset.py:160-163,203-206), everything else happens in `collate` for the whole batch on the
    device.  `clips`: list of dicts {"audio": 1-D float array at `sr`, "bitstream": '0101..' ('0' = silent)}; `noises`: list of 1-D
    float arrays, each at least as long as a clip; `snrs` / `snr_idx` as in the reference (random choice per item when snr_idx is
    None, M2/dataset.py:198-201).  Clips of one dataset have equal length (the reference slices `data_len_sec`)."""

    def __init__(self, clips, noises, sr=16000, fps=30.0, snrs=SNRS, snr_idx=None, norm=0.5, seed=None):
        if not clips:
            raise ValueError("WaveformDataset needs at least one clip")
        n = len(clips[0]["audio"])
        for c in clips:
            if len(c["audio"]) != n:
                raise RuntimeError("clips of one dataset must have equal length")        # (dataset errors are RuntimeError, M2/dataset.py:307-309)
            if set(c["bitstream"]) - {"0", "1"}:
                raise RuntimeError("Invalid bit?")
        self.clips, self.noises = clips, [torch.as_tensor(x, dtype=torch.float32) for x in noises]
        self.sr, self.fps, self.snrs, self.snr_idx, self.norm = sr, fps, list(snrs), snr_idx, norm
        self.rng = random.Random(seed)

    def __len__(self):
        return len(self.clips)

    def __getitem__(self, i):
        c = self.clips[i]
        n = len(c["audio"])
        noise = self.rng.choice(self.noises)
        if noise.numel() < n:
            raise RuntimeError(f"noise track of {noise.numel()} samples is shorter than the {n}-sample clip")
        start = self.rng.randint(0, noise.numel() - n)
        snr = self.rng.choice(self.snrs) if self.snr_idx is None else self.snrs[self.snr_idx]
        return {"audio": torch.as_tensor(c["audio"], dtype=torch.float32), "noise": noise[start:start + n], "snr": float(snr),
                "bitstream": c["bitstream"], "start": 0}

    def collate(self, batch, device=None, model="joint"):
        """list of raw items -> the reference's item dict for the whole batch, built on the device (pinned staging, async copies)."""
        device = device or torch.device("cuda", torch.cuda.current_device())
        audio = torch.stack([b["audio"] for b in batch]).pin_memory().to(device, non_blocking=True)
        noise = torch.stack([b["noise"] for b in batch]).pin_memory().to(device, non_blocking=True)
        snr = [b["snr"] for b in batch]
        bits = [b["bitstream"] for b in batch]
        if model == "sid":
            return make_sid_items(audio, noise, snr, bits, self.sr, self.fps, self.norm)
        item = make_joint_items(audio, noise, snr, bits, self.sr, self.fps, self.norm)
        item["bitstream"] = bits
        item["start"] = [b["start"] for b in batch]
        return item
