#!/usr/bin/env python
"""Benchmark of the two-stage denoising hot path (BASELINE.json metric: clips/sec, 2 s @ 16 kHz, fwd+bwd).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path, one rank per GPU (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port) on the host cores

One "step" = BASELINE.json configs[1]: a batch of 32 synthetic 2 s clips per GPU through
  4 x STFT (mixed, silent-interval-gated noise, clean, full noise)  ->  SID fwd+bwd (BCE)  ->  JointModel fwd+bwd (2 x MSE
  through the cRM recovery)  ->  iSTFT of the recovered spectrogram  ->  fused Adam on both networks
(+ one NCCL all-reduce of each flat gradient buffer when N > 1; per-GPU batch fixed = weak scaling).
`value` is measured with the waveforms already resident in HBM; `e2e` repeats the measurement through the same public API
with the step's inputs coming from pinned host memory and the losses + denoised waveforms read back every step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

SR, LENGTH, FPS = 16000, 32000, 30.0
SNRS = [-10, -7, -3, 0, 3, 7, 10]
METRIC = "clips/sec (2 s @16 kHz) fwd+bwd"
WORKLOAD = "batch=32 2 s clips per GPU, full STFT->SID->U-Net->iSTFT fwd+bwd training step (BASELINE configs[1])"


# ----------------------------------------------------------------------------------------------- synthetic clips
def synth_batch(batch, length=LENGTH, sr=SR, start=0):
    """Deterministic noisy-speech clips (SURVEY.md 8d recipe, vectorised): harmonic 'speech' silenced on random silent
    intervals, low-passed Gaussian noise, mixed at SNRS[i % 7] and normalised to max |mixed| = 0.5."""
    from scipy.signal import lfilter
    n_bits = int(round(length / sr * FPS))
    ratio = sr / FPS
    t = np.arange(length) / sr
    out = {k: np.zeros((batch, length), np.float32) for k in ("mixed", "clean", "full_noise")}
    bits = np.zeros((batch, n_bits), np.uint8)
    for i in range(batch):
        rng = np.random.default_rng(1234 + start + i)
        b, state = [], int(rng.integers(0, 2))
        while len(b) < n_bits:
            b.extend([state] * int(rng.integers(3, 21)))
            state ^= 1
        b = np.array(b[:n_bits], np.uint8)
        bits[i] = b
        mask = np.zeros(length)
        for j in range(n_bits):
            if b[j] == 0:
                lo, hi = int(j * ratio), int((j + 1) * ratio - 1)
                mask[lo:hi + (1 if j + 1 < n_bits and b[j + 1] == 0 else 0)] = 1
        f0 = rng.uniform(100, 250)
        speech = sum(np.sin(2 * np.pi * f0 * h * t + rng.uniform(0, 2 * np.pi)) / h for h in range(1, 6))
        speech = speech * 0.5 * (1 - np.cos(2 * np.pi * 4.0 * t)) * (1 - mask)
        noise = lfilter([0.05], [1, -0.95], rng.standard_normal(length))
        sp, snr = np.sum(speech ** 2), SNRS[(start + i) % 7]
        if sp > 0:
            noise = noise * np.sqrt(sp / 10 ** (snr / 10)) / (np.sqrt(np.sum(noise ** 2)) + 1e-30)
        mixed = speech + noise
        scale = np.max(np.abs(mixed)) / 0.5
        out["mixed"][i], out["clean"][i], out["full_noise"][i] = mixed / scale, speech / scale, noise / scale
    out["bits"] = bits
    out["label"] = bits.astype(np.float32)
    return out


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler(threading.Thread):
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.begin = index, [], False, 0

    def mark_begin(self):
        """Only samples taken from now on count (the sampler itself is started BEFORE the warm-up steps: spawning nvidia-smi from
        a large process holds the interpreter for ~0.1 s, which must not fall into the timed region)."""
        self.begin = len(self.samples)

    def run(self):
        try:
            p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits",
                                  "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.proc = p
        for line in p.stdout:
            if self.stop_flag:
                break
            self.samples.append([c.strip() for c in line.split(",")])
        p.kill()

    def summary(self):
        self.stop_flag = True
        if hasattr(self, "proc"):
            self.proc.kill()
        self.samples = self.samples[self.begin:] or self.samples
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(len(s) > 3 + j and s[3 + j].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.samples[0][1]), "power_w_max": max(float(s[2]) for s in self.samples),
                "samples": len(sm), "reasons": reasons}


# ----------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_step_factory(batch, length=LENGTH):
    """The reference algorithm on the host cores through the CPU oracle (a port: /root/reference does not travel to the
    GPU box): numpy/torch STFT x4 -> SID fwd+bwd -> Joint fwd+bwd -> iSTFT -> Adam.  Returns (step_fn, n_clips)."""
    import torch
    import torch.nn.functional as F
    from oracle import nets, transform as otf
    from oracle.gating import bits_to_sample_mask
    torch.set_num_threads(os.cpu_count() or 1)
    data = synth_batch(batch, length)
    sid_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in nets.synth_state_dict(nets.sid_shapes(), 3).items()}
    jt_sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in nets.synth_state_dict(nets.joint_shapes(), 4).items()}
    params = [v for v in list(sid_sd.values()) + list(jt_sd.values()) if v.requires_grad]
    opt = torch.optim.Adam(params, 1e-3)
    ratio = SR / FPS
    bit_strings = ["".join(str(int(b)) for b in row) for row in data["bits"]]

    def step():
        masks = np.stack([bits_to_sample_mask(length, ratio, b) for b in bit_strings]).astype(np.float32)
        mixed = torch.tensor(otf.stft_batch(data["mixed"]))
        noise = torch.tensor(otf.stft_batch(data["mixed"] * masks))
        clean = torch.tensor(otf.stft_batch(data["clean"]))
        full = torch.tensor(otf.stft_batch(data["full_noise"]))
        opt.zero_grad()
        logits = nets.sid_forward(sid_sd, mixed, data["label"].shape[1], training=True)
        l0 = F.binary_cross_entropy_with_logits(logits, torch.tensor(data["label"]))
        l0.backward()
        n_pred, mask = nets.joint_forward(jt_sd, mixed, noise, training=True)
        rec = otf.batch_fast_icRM_sigmoid(mixed, mask)
        l1, l2 = F.mse_loss(n_pred, full), F.mse_loss(rec, clean)
        (l1 + l2).backward()
        opt.step()
        otf.istft_batch(rec.detach().numpy())
        return float(l0.detach()), float(l1.detach()), float(l2.detach())

    return step, batch


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.ref_batch
    step, n = cpu_step_factory(batch)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    cores = os.cpu_count() or 1
    sample = f"{batch} clips per step (same STFT->SID->Joint->iSTFT->Adam fwd+bwd step, BatchNorm over {batch} clips), {args.steps} steps"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# ----------------------------------------------------------------------------------------------- this repo's arm (GPU)
def cpu_forward_b1():
    """BASELINE configs[0]: ONE 2 s clip, SID + JointModel forward only, eval mode, on the host cores (oracle port)."""
    import torch
    from oracle import nets, transform as otf
    torch.set_num_threads(os.cpu_count() or 1)
    data = synth_batch(1)
    sid_sd, jt_sd = nets.synth_state_dict(nets.sid_shapes(), 3), nets.synth_state_dict(nets.joint_shapes(), 4)
    ratio = SR / FPS

    def fwd():
        with torch.no_grad():
            mixed = torch.tensor(np.ascontiguousarray(otf.stft_batch(data["mixed"])))
            logits = nets.sid_forward(sid_sd, mixed, data["label"].shape[1], training=False)
            bits = (torch.sigmoid(logits) >= 0.5).numpy().astype(np.uint8)
            from oracle.gating import bits_to_sample_mask
            mask = bits_to_sample_mask(LENGTH, ratio, "".join(str(int(b)) for b in bits[0])).astype(np.float32)
            noise = torch.tensor(np.ascontiguousarray(otf.stft_batch(data["mixed"] * mask[None])))
            _, m = nets.joint_forward(jt_sd, mixed, noise, training=False)
            rec = otf.batch_fast_icRM_sigmoid(mixed, m)
            otf.istft_batch(rec.numpy())
    fwd()
    t0 = time.perf_counter()
    for _ in range(3):
        fwd()
    return (time.perf_counter() - t0) / 3


def run_gpu(args):
    import torch
    import torch.distributed as dist
    import sos_b200
    from sos_b200 import _lib, agent as ag, ops, tools, transform

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sos_b200 hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ops.init()
    B = args.batch
    ratio = SR / FPS

    torch.manual_seed(0)
    sid = ag.SIDAgent(ag.default_config(model="sid"))
    torch.manual_seed(1)
    joint = ag.MyAgent(ag.default_config(model="joint", sr=SR, fps=FPS))

    host = synth_batch(B, LENGTH, SR, start=rank * B)
    pinned = {k: torch.from_numpy(v).pin_memory() for k, v in host.items()}
    resident = {k: v.to(dev) for k, v in pinned.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values())
    out_host = {"wave": torch.empty(B, 158 * (LENGTH // 158), dtype=torch.float32).pin_memory(), "loss": torch.empty(3, dtype=torch.float32).pin_memory()}
    d2h_bytes = sum(v.numel() * v.element_size() for v in out_host.values())
    trainer = ag.GraphedTrainStep(sid, joint, B, LENGTH, SR, FPS, warmup=2) if args.graph else None

    def eager_step(d):
        # the four transforms of a training item (M2/dataset.py:234-237) as ONE launch over the 4 x B waveforms
        gated = tools.gate_noise(d["mixed"], ratio, d["bits"])
        spec = transform.stft_batch(torch.cat([d["mixed"], gated, d["clean"], d["full_noise"]]))
        mixed, noise, clean, full = spec[:B], spec[B:2 * B], spec[2 * B:3 * B], spec[3 * B:]
        _, l_sid = sid.train_func({"audio": mixed, "label": d["label"]})
        _, l_jt = joint.train_func({"mixed": mixed, "noise": noise, "clean": clean, "full_noise": full})
        wave = transform.istft_batch(joint.last_rec.detach())
        return {"losses": torch.stack([l_sid["bce"].detach(), l_jt["stage1"].detach(), l_jt["stage2"].detach()]), "wave": wave}

    def step(src, e2e):
        d = {k: v.to(dev, non_blocking=True) for k, v in src.items()} if e2e else src
        out = trainer(d["mixed"], d["clean"], d["full_noise"], d["bits"], d["label"]) if trainer else eager_step(d)
        if e2e:
            out_host["loss"].copy_(out["losses"], non_blocking=True)
            out_host["wave"].copy_(out["wave"], non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the user holds the step's result before the next step
        return out

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(src, e2e, steps):
        import gc
        gc.collect()
        gc.disable()                  # no collector pauses while the host feeds the GPU
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h0 = time.perf_counter()
        for _ in range(steps):
            step(src, e2e)
        e1.record()
        host_ms[0] = 1e3 * (time.perf_counter() - h0) / steps          # host time to ENQUEUE a step (no synchronisation inside)
        barrier()
        gc.enable()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    host_ms = [0.0]

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = _lib.launch_count
    for _ in range(max(args.warmup, 3) + (1 if trainer else 0)):       # (graph mode: 2 eager steps, the capture, then replays)
        step(resident, False)
    warm_launches = _lib.launch_count - n0
    if sampler:
        sampler.mark_begin()
    # ---- the metric: NO per-launch instrumentation inside this region
    n0 = _lib.launch_count
    ms = timed(resident, False, args.steps)
    enqueue_ms = host_ms[0]
    launches = _lib.launch_count - n0
    if trainer:                       # replays do not pass through the Python binding: count what the captured graphs hold
        launches = (trainer.launches_per_step + 4) * args.steps
    clocks = sampler.summary() if sampler else None
    step(pinned, True)
    ms_e2e = timed(pinned, True, args.steps)
    clips = world * B * args.steps

    # ---- per-kernel breakdown in a SEPARATE eager pass (CUDA-event pair around every tensor-core / transform launch, weight
    #      gradients on the main stream so that an event pair brackets exactly one kernel)
    os.environ["SOS_SYNC_WGRAD"] = "1"
    eager_step(resident)
    ops.profile_start()
    prof_steps = 2
    for _ in range(prof_steps):
        eager_step(resident)
    prof = ops.profile_stop()
    del os.environ["SOS_SYNC_WGRAD"]
    barrier()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        bf16_peak = peaks.get("bf16_tflops_sustained", 1400.0)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        peak_src = "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)" if peaks else "fallback (B200_PROFILING.md)"
        kern = {}
        for name, rec in prof.items():
            if rec["ms"] > 0:
                kern[name] = {"launches_per_step": rec["n"] / prof_steps, "ms_per_step": rec["ms"] / prof_steps, "avg_launch_us": 1e3 * rec["ms"] / max(rec["n"], 1),
                              "tflops": rec["flops"] / (rec["ms"] * 1e-3) / 1e12 if rec["flops"] else None,
                              "gbs": rec["bytes"] / (rec["ms"] * 1e-3) / 1e9 if rec["bytes"] else None}
        for name in ("stft", "istft"):
            if name in kern and kern[name]["gbs"]:
                kern[name]["hbm_frac"] = kern[name]["gbs"] / hbm_peak
        # the dominant kernel: tapgemm_f16_kernel = every convolution's forward AND data gradient (same kernel, flipped taps).
        # roofline = ALL its launches of a step (algorithmic FLOPs / summed launch durations)
        ncu = {}
        try:
            ncu = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        except (OSError, ValueError):
            pass
        recs = [prof[n] for n in ("conv_fwd", "conv_dgrad") if n in prof and prof[n]["ms"] > 0]
        # the other tensor-bound kernel families of the step, same accounting (algorithmic FLOPs of their launches / summed durations)
        fam = {"rowconv_f16_kernel (48-channel dilated 5x5 layers: forward + data gradient)": ("conv_fwd_row", "conv_dgrad_row"),
               "wgrad_f16_kernel (all weight gradients)": ("conv_wgrad",)}
        other = []
        for label, names in fam.items():
            rr = [prof[n] for n in names if n in prof and prof[n]["ms"] > 0]
            if rr:
                fl, tms, nl = sum(r["flops"] for r in rr), sum(r["ms"] for r in rr), sum(r["n"] for r in rr)
                tr = ncu.get(label.split(" ")[0], {})
                other.append({"bound": "tensor", "kernel": label, "achieved": fl / (tms * 1e-3) / 1e12, "peak": bf16_peak, "unit": "TFLOP/s",
                              "frac": fl / (tms * 1e-3) / 1e12 / bf16_peak, "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note"),
                              "launches_per_step": nl / prof_steps, "ms_per_step": tms / prof_steps})
        roofline = None
        if recs:
            fl, tms, nl = sum(r["flops"] for r in recs), sum(r["ms"] for r in recs), sum(r["n"] for r in recs)
            ach = fl / (tms * 1e-3) / 1e12
            tr = ncu.get("tapgemm_f16_pair_kernel", ncu.get("tapgemm_f16_kernel", {}))
            roofline = {"bound": "tensor", "kernel": "tapgemm_f16_pair_kernel / tapgemm_f16_kernel (all launches of a step: conv forward + data gradient "
                                                     "of every layer the row-streaming kernel does not serve)",
                        "achieved": ach, "peak": bf16_peak, "unit": "TFLOP/s", "frac": ach / bf16_peak,
                        "traffic": tr.get("dram_bytes_per_launch"), "traffic_note": tr.get("note"),
                        "launches_per_step": nl / prof_steps, "avg_launch_us": 1e3 * tms / nl, "flops_per_launch": fl / nl,
                        "peak_source": peak_src + "; the kernel issues tcgen05 kind::f16 (same rate as bf16)",
                        "ms_per_step": tms / prof_steps, "share_of_step": (tms / prof_steps) / (ms / args.steps),
                        "measured": "CUDA events around every launch in a separate eager pass (not inside the timed `value` region)"}
        # whole-step conv fraction: ALL algorithmic conv FLOPs of the step (fwd + dgrad + wgrad = 3 x 506.45 GFLOP per clip,
        # SURVEY 8d) over the step time, against the same peak
        conv_tflop_per_clip = 3 * 506.45e-3
        step_conv_frac = B * conv_tflop_per_clip / (ms / args.steps * 1e-3) / bf16_peak
        cpu, extra = None, None
        if world == 1 and not args.no_cpu_baseline:
            cstep, n = cpu_step_factory(args.ref_batch)
            cstep()                                        # warm-up (thread pools, allocator)
            t0 = time.perf_counter()
            reps = 2
            for _ in range(reps):
                cstep()
            dt = (time.perf_counter() - t0) / reps
            cpu = {"value": n / dt, "unit": "clips/s", "cores": os.cpu_count() or 1, "kind": "port", "same_config": n == B,
                   "sample": f"{n}-clip steps of the same workload (BatchNorm over {n} clips; {B} clips would need ~2 GB of host memory per clip "
                             f"for the fp32 autograd graph), 1 warm-up + {reps} timed, {dt:.2f} s per step"}
        if world == 1 and not args.no_extra:
            extra = run_extra(dev, sid.net, joint.net, not args.no_cpu_baseline)
        print(json.dumps({
            "metric": METRIC, "value": clips / (ms * 1e-3), "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 conv operands (11-bit significand, as TF32), fp32 accumulate; fp32 elsewhere",
            "data": "synthetic", "host_enqueue_ms_per_step": enqueue_ms,
            "config": {"workload": WORKLOAD, "per_gpu_batch": B, "global_batch": world * B, "samples_per_clip": LENGTH, "frames": 1 + LENGTH // 158,
                       "parallelism": f"dp{world} (NCCL all-reduce of the flat gradient buffers, SID's overlapped with the Joint step)" if world > 1 else "single GPU",
                       "cuda_graph": bool(trainer),
                       "l2": "no explicit flush: the step's activation working set (tens of GB) is far larger than the 126 MB L2"},
            "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": "clips/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_other_kernels": other, "step_conv_frac": step_conv_frac,
            "step_conv_frac_note": f"{B} clips x 1.519 TFLOP (conv fwd + dgrad + wgrad) / ms_per_step / {bf16_peak:.1f} TFLOP/s",
            "kernels": kern, "cpu_baseline": cpu, "extra": extra}))
    if world > 1:
        dist.destroy_process_group()


def run_extra(dev, sid_net, joint_net, with_cpu):
    """The other BASELINE configs, measured in the same run (parity-test cases, not the headline): configs[0] one clip forward
    (GPU latency + the CPU port's), configs[3] forward-only sweep, configs[4] 10 s clips x 16 with chunked STFT / streaming OLA."""
    import torch
    from sos_b200 import pipeline
    sid_net.eval()
    joint_net.eval()
    res = {"infer_sweep": []}

    def run(B, L, fpc=None, steps=5, graph=True):
        wave = torch.from_numpy(synth_batch(min(B, 32), L)["mixed"]).repeat((B + 31) // 32, 1)[:B].contiguous().to(dev)
        fn = pipeline.GraphedDenoiser(sid_net, joint_net, B, L, SR, FPS, frames_per_chunk=fpc) if graph else \
            (lambda w: pipeline.denoise(w, sid_net, joint_net, SR, FPS, frames_per_chunk=fpc))
        for _ in range(3):
            fn(wave)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn(wave)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del fn
        torch.cuda.empty_cache()
        return {"batch": B, "samples_per_clip": L, "ms_per_batch": ms, "clips_per_s": B / (ms * 1e-3), "cuda_graph": graph}

    with torch.no_grad():
        for B in (1, 8, 32, 128, 512):
            res["infer_sweep"].append(run(B, LENGTH, steps=10 if B <= 32 else 3))
        res["config0_single_clip_forward"] = dict(res["infer_sweep"][0])
        res["config4_longform_10s_x16"] = run(16, 160000, fpc=256, steps=3)
    if with_cpu:
        dt = cpu_forward_b1()
        res["config0_single_clip_forward"]["cpu_port_ms"] = 1e3 * dt
        res["config0_single_clip_forward"]["cpu_cores"] = os.cpu_count() or 1
    sid_net.train()
    joint_net.train()
    return res


def run_infer(args):
    """BASELINE configs[0] / configs[3]: forward-only in-memory inference (STFT -> SID -> gate -> STFT -> JointModel -> cRM + iSTFT)
    at --batch clips; NOT the default workload (that is the training step).  One JSON line."""
    import torch
    import sos_b200  # noqa: F401
    from sos_b200 import _lib, networks, ops, pipeline
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the sos_b200 hot path has no CPU fallback")
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    ops.init()
    B, L = args.batch, args.length
    torch.manual_seed(0)
    sid = networks.get_network().to(dev).eval()
    torch.manual_seed(1)
    joint = networks.get_network(object()).to(dev).eval()
    host = torch.from_numpy(synth_batch(min(B, 32), L)["mixed"]).repeat((B + 31) // 32, 1)[:B].contiguous().pin_memory()
    resident = host.to(dev)
    out_host = torch.empty(B, 158 * (L // 158), dtype=torch.float32).pin_memory()
    fpc = 256 if L > 64000 else None
    graphed = pipeline.GraphedDenoiser(sid, joint, B, L, SR, FPS, frames_per_chunk=fpc) if args.graph else None

    def step(e2e):
        w = host.to(dev, non_blocking=True) if e2e else resident
        out = graphed(w) if graphed else pipeline.denoise(w, sid, joint, SR, FPS, frames_per_chunk=fpc)
        if e2e:
            out_host.copy_(out["denoised"], non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def timed(e2e):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(e2e)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for _ in range(max(args.warmup, 3)):
        step(False)
    n0 = _lib.launch_count
    ms = timed(False)
    launches = _lib.launch_count - n0
    ms_e2e = timed(True)
    clips = B * args.steps
    print(json.dumps({
        "metric": "clips/sec forward-only inference", "value": clips / (ms * 1e-3), "unit": "clips/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 conv operands (11-bit significand, as TF32), fp32 accumulate; fp32 elsewhere", "data": "synthetic",
        "config": {"workload": f"batch={B} clips of {L} samples @16 kHz, STFT->SID->gate->STFT->JointModel->cRM+iSTFT, forward only "
                               "(BASELINE configs[0]/[3]/[4])", "batch": B, "samples_per_clip": L, "frames": 1 + L // 158,
                   "chunked_transforms": bool(fpc), "cuda_graph": bool(graphed)},
        "e2e": {"value": clips / (ms_e2e * 1e-3), "unit": "clips/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": launches}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="sos_b200", choices=["sos_b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU")
    ap.add_argument("--ref-batch", type=int, default=8, help="clips per step of the CPU arm (a bounded sample of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the configs[0]/[3]/[4] forward-only measurements")
    ap.add_argument("--workload", default="train", choices=["train", "infer"], help="train = BASELINE configs[1] (default, the metric); "
                    "infer = forward-only inference at --batch / --length")
    ap.add_argument("--length", type=int, default=LENGTH, help="samples per clip (infer workload)")
    ap.add_argument("--graph", type=int, default=1, help="replay the step as CUDA graphs (train: agent.GraphedTrainStep; infer: pipeline.GraphedDenoiser)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "infer":
        run_infer(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
