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
