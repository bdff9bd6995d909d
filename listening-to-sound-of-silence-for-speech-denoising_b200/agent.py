"""Training-step drivers with the reference's agent surface, on the sm_100a kernels.

  get_agent(config)                      M2/agent.py:16-17, M1/agent.py:16-17
  MyAgent.forward(data)                  M2/agent.py:176-190   -> ((n_pred, mask), {"stage1", "stage2"})
  SIDAgent.forward(data)                 M1/agent.py:189-206   -> (logits, {"bce"})
  update_network / train_func / val_func M2/agent.py:101-141
  save_ckpt / load_ckpt                  M2/agent.py:57-95     (same dict keys; model_state_dict has the reference's names)
  build_net (nn.DataParallel)            M2/agent.py:153-165   -> replaced by one process per GPU + ONE NCCL all-reduce of the
                                                                  flat gradient buffer per step (FlatAdam.step)

Differences from the reference that are part of the design: the optimiser state lives in one flat fp32 buffer
(parameters, gradients, exp_avg, exp_avg_sq are views of four contiguous allocations) so that Adam is one kernel launch
and the data-parallel exchange is one collective; the data dict may carry WAVEFORMS ("mixed_wave", ... + "bits"), in which
case the four STFTs and the silent-interval gate run on the device (the reference does them in DataLoader workers).
"""
import os

import torch
import torch.nn as nn

from . import layers as L
from . import ops, tools, transform
from .networks import get_network

PHASE_TRAINING, PHASE_TESTING = "train", "test"


class TrainClock(object):
    """M2/utils.py TrainClock: epoch / minibatch / step counters with the same checkpoint dict."""

    def __init__(self):
        self.epoch, self.minibatch, self.step = 1, 0, 0

    def tick(self):
        self.minibatch += 1
        self.step += 1

    def tock(self):
        self.epoch += 1
        self.minibatch = 0

    def make_checkpoint(self):
        return {"epoch": self.epoch, "minibatch": self.minibatch, "step": self.step}

    def restore_checkpoint(self, d):
        self.epoch, self.minibatch, self.step = d["epoch"], d["minibatch"], d["step"]


class FlatAdam(object):
    """optim.Adam(params, lr) (betas 0.9/0.999, eps 1e-8, no weight decay; M2/agent.py:167-170) over ONE flat buffer.

    Parameters and their .grad become views of two contiguous fp32 allocations, so a step is: (optional) one NCCL
    all-reduce of the flat gradient (sum, then 1/world folded into the Adam kernel's grad_scale) + one Adam kernel.
    `param_groups` / `state_dict()` keep the shape torch's StepLR and the reference's checkpoint code expect.
    """

    def __init__(self, params, lr, betas=(0.9, 0.999), eps=1e-8, process_group=None):
        self.params = [p for p in params if p.requires_grad]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        # keep every slice 16-byte aligned for the vectorised kernels
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 3) // 4 * 4
        self.numel = off
        self.flat_param = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(off, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(off, device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            self.flat_param[o:o + p.numel()].copy_(p.data.reshape(-1))
            p.data = self.flat_param[o:o + p.numel()].view(p.shape)
            p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)
        self.param_groups = [{"lr": lr, "initial_lr": lr, "betas": betas, "eps": eps, "params": self.params}]
        self.defaults = {"lr": lr}
        self.step_count = 0
        self.group = process_group
        self.n_real = n
        # the optimiser clock lives on the device (sos_adam_step_dev): [lr, step, lr / (1 - b1^step), 1 / sqrt(1 - b2^step)], so a
        # captured CUDA graph of the step replays with the right bias correction; the host mirrors step_count for checkpoints
        self.state = torch.tensor([lr, 0.0, 0.0, 0.0], device=dev, dtype=torch.float32)
        self._state_lr = lr
        self.broadcast()

    def broadcast(self):
        """Data parallelism needs bit-identical replicas: rank 0's parameters (and optimiser moments) overwrite every other
        rank's, at construction and after a checkpoint is loaded (no reliance on identical seeds)."""
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size(self.group) > 1:
            for t in (self.flat_param, self.exp_avg, self.exp_avg_sq, self.state):
                torch.distributed.broadcast(t, 0, group=self.group)

    def sync_clock(self):
        """Push a host-side change of the learning rate (StepLR) or of the step count (checkpoint restore) to the device clock.
        Called OUTSIDE captured regions; a no-op (no launch) when nothing changed."""
        lr = self.param_groups[0]["lr"]
        if lr != self._state_lr:
            self.state[0:1].fill_(lr)
            self._state_lr = lr

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        self.zero_grad_views()

    def zero_grad_views(self):
        for p, o in zip(self.params, self.offsets):            # autograd may have replaced .grad; re-attach the views
            if p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * o:
                p.grad = self.flat_grad[o:o + p.numel()].view(p.shape)

    def exchange(self):
        """The data-parallel exchange: ONE all-reduce (sum) of the flat gradient buffer (NCCL over NVLink on the GPUs; the
        CPU tests drive the same code over gloo).  Returns the world size; the 1/world factor is folded into the Adam kernel."""
        world = 1
        if self.group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            world = torch.distributed.get_world_size(self.group)
            if world > 1:
                torch.distributed.all_reduce(self.flat_grad, group=self.group)        # ncclAllReduce(sum) over NVLink
        return world

    def step(self, exchange=True, world=None):
        """exchange=False: the caller has already all-reduced the flat gradient (or captured this call in a CUDA graph, where
        the collective stays outside) and passes the world size."""
        if exchange:
            world = self.exchange()
        elif world is None:
            world = 1
        self.sync_clock()
        self.apply(world)

    def apply(self, world=1):
        """The update itself: two launches (clock tick + fused Adam over the flat buffers); graph-capturable."""
        self.step_count += 1
        g = self.param_groups[0]
        ops.adam_step_dev(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, self.state, g["betas"][0], g["betas"][1], g["eps"],
                          grad_scale=1.0 / world)

    # torch.optim-shaped checkpoint: {"state": {i: {step, exp_avg, exp_avg_sq}}, "param_groups": [...]}
    def state_dict(self):
        state = {}
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            state[i] = {"step": torch.tensor(float(self.step_count)),
                        "exp_avg": self.exp_avg[o:o + p.numel()].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + p.numel()].view(p.shape).clone()}
        g = self.param_groups[0]
        return {"state": state if self.step_count else {},
                "param_groups": [{"lr": g["lr"], "betas": g["betas"], "eps": g["eps"], "weight_decay": 0, "amsgrad": False,
                                  "initial_lr": g["initial_lr"], "params": list(range(len(self.params)))}]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.param_groups[0].update(lr=g["lr"], betas=tuple(g["betas"]), eps=g["eps"])
        self.param_groups[0]["initial_lr"] = g.get("initial_lr", g["lr"])
        for i, (p, o) in enumerate(zip(self.params, self.offsets)):
            st = sd["state"].get(i)
            if st is None:
                continue
            self.step_count = int(float(st["step"]))
            self.exp_avg[o:o + p.numel()].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + p.numel()].copy_(st["exp_avg_sq"].reshape(-1))
        self.state[0:1].fill_(self.param_groups[0]["lr"])
        self.state[1:2].fill_(float(self.step_count))
        self._state_lr = self.param_groups[0]["lr"]


class StepLR(object):
    """optim.lr_scheduler.StepLR(optimizer, step_size, gamma=0.1) stepped once per epoch (M2/agent.py:108-111,170)."""

    def __init__(self, optimizer, step_size, gamma=0.1):
        self.optimizer, self.step_size, self.gamma, self.last_epoch = optimizer, step_size, gamma, 0

    def step(self, epoch=None):
        self.last_epoch = self.last_epoch + 1 if epoch is None else epoch
        for g in self.optimizer.param_groups:
            g["lr"] = g["initial_lr"] * self.gamma ** (self.last_epoch // self.step_size)

    def state_dict(self):
        return {"step_size": self.step_size, "gamma": self.gamma, "last_epoch": self.last_epoch}

    def load_state_dict(self, d):
        self.step_size, self.gamma, self.last_epoch = d["step_size"], d["gamma"], d["last_epoch"]


class _Config(object):
    """The attributes of M2/common.py Config / M1/common.py Config that the agents read."""
    lr = 1e-3                   # M2/common.py:55
    lr_step_size = 15           # M2/common.py:56
    batch_size = 40
    model = "joint"             # "joint" (M2) or "sid" (M1)
    log_dir = None
    model_dir = None
    sr = 16000
    fps = 30.0


def default_config(**kw):
    c = _Config()
    for k, v in kw.items():
        setattr(c, k, v)
    return c


class BaseAgent(object):
    def __init__(self, config):
        ops.init()
        self.config = config
        self.log_dir = getattr(config, "log_dir", None)
        self.model_dir = getattr(config, "model_dir", None)
        self.clock = TrainClock()
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.batch_size = getattr(config, "batch_size", None)
        self.net = self.build_net(config)
        self.set_loss_function()
        self.set_optimizer(config)
        self.last_losses = {}
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            for b in self.net.buffers():                        # BatchNorm running statistics start identical on every rank as well
                torch.distributed.broadcast(b, 0)

    # -- reference: nn.DataParallel(net.cuda()) when several GPUs are visible.  Here: this process owns ONE GPU; when
    #    torch.distributed is initialised every rank builds the same net (same seed) and FlatAdam all-reduces gradients.
    def build_net(self, config):
        raise NotImplementedError

    def set_loss_function(self):
        self.criterion = L.MSELoss.apply

    def set_optimizer(self, config):
        self.optimizer = FlatAdam(self.net.parameters(), getattr(config, "lr", 1e-3))
        self.scheduler = StepLR(self.optimizer, getattr(config, "lr_step_size", 15))

    def save_ckpt(self, name=None):
        path = os.path.join(self.model_dir, ("ckpt_epoch{}.pth".format(self.clock.epoch)) if name is None else "{}.pth".format(name))
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_rank() != 0:
            return path                                         # replicas are identical: rank 0 writes the file
        torch.save({"clock": self.clock.make_checkpoint(),
                    "model_state_dict": {k: v.detach().cpu().clone() for k, v in self.net.state_dict().items()},
                    "optimizer_state_dict": self.optimizer.state_dict(),
                    "scheduler_state_dict": self.scheduler.state_dict()}, path)
        return path

    def load_ckpt(self, name=None):
        name = name if name == "latest" else "ckpt_epoch{}".format(name)
        path = os.path.join(self.model_dir, "{}.pth".format(name))
        if not os.path.exists(path):
            raise ValueError("Checkpoint {} not exists.".format(path))
        ck = torch.load(path, map_location="cpu")
        with torch.no_grad():                                   # copy INTO the flat views (keeps the optimiser's aliasing)
            own = self.net.state_dict()
            missing, unexpected = sorted(set(own) - set(ck["model_state_dict"])), sorted(set(ck["model_state_dict"]) - set(own))
            if missing or unexpected:                           # strict, like the reference's load_state_dict (M2/agent.py:89)
                raise RuntimeError("Error(s) in loading state_dict: missing keys {}, unexpected keys {}".format(missing, unexpected))
            for k, v in ck["model_state_dict"].items():
                own[k].copy_(v)
        L.weights_changed()
        self.optimizer.load_state_dict(ck["optimizer_state_dict"])
        self.optimizer.broadcast()
        self.scheduler.load_state_dict(ck["scheduler_state_dict"])
        self.clock.restore_checkpoint(ck["clock"])

    def forward(self, data):
        raise NotImplementedError

    def backward(self, loss_dict):
        """zero_grad + backward of the summed losses (M2/agent.py:101-105) into the flat gradient buffer; graph-capturable."""
        loss = sum(loss_dict.values())
        self.optimizer.zero_grad()
        if os.environ.get("SOS_SYNC_WGRAD"):
            loss.backward()
        else:
            with L.async_wgrad():                              # conv weight gradients on a side stream, joined on exit
                loss.backward()

    def update_network(self, loss_dict):
        self.backward(loss_dict)
        self.optimizer.step()
        L.weights_changed()

    def update_learning_rate(self):
        self.scheduler.step(self.clock.epoch)

    def record_losses(self, loss_dict, mode=PHASE_TRAINING):
        # the reference calls .item() on every loss here (a device sync per loss per step, M2/agent.py:113-119); the
        # losses stay on the device and are read only when someone asks (`loss_values()`).
        self.last_losses = {k: v.detach() for k, v in loss_dict.items()}

    def loss_values(self):
        return {k: float(v) for k, v in self.last_losses.items()}

    def train_func(self, data):
        self.net.train()
        L.pack_all()                                             # every weight operand the optimiser step invalidated, one launch
        outputs, losses = self.forward(data)
        self.update_network(losses)
        self.record_losses(losses, PHASE_TRAINING)
        return outputs, losses

    def val_func(self, data):
        self.net.eval()
        with torch.no_grad():
            outputs, losses = self.forward(data)
        self.record_losses(losses, PHASE_TESTING)
        return outputs, losses


def _dev(t, device):
    return t.to(device, non_blocking=True) if torch.is_tensor(t) else torch.as_tensor(t).to(device, non_blocking=True)


class MyAgent(BaseAgent):
    """Stage-2 (JointModel) trainer, M2/agent.py:144-233."""

    def build_net(self, config):
        return get_network(config).to(self.device)

    def spectrograms(self, data):
        """Dataset item dict (M2/dataset.py:311-320) with spectrograms, or waveforms + bit strings to transform here."""
        d = self.device
        if "mixed" in data:
            return tuple(_dev(data[k], d).float() for k in ("mixed", "noise", "clean", "full_noise"))
        mixed_w = _dev(data["mixed_wave"], d)
        bits = _dev(data["bits"], d)
        ratio = getattr(self.config, "sr", 16000) / getattr(self.config, "fps", 30.0)
        # the item's four transforms (M2/dataset.py:234-237) as ONE launch over the 4 x B waveforms: more than one wave of CTAs
        B = mixed_w.shape[0]
        gated = tools.gate_noise(mixed_w, ratio, bits)                        # noise_sig = mixed * mask  (M2/dataset.py:229)
        spec = transform.stft_batch(torch.cat([mixed_w, gated, _dev(data["clean_wave"], d), _dev(data["full_noise_wave"], d)]))
        return spec[:B], spec[B:2 * B], spec[2 * B:3 * B], spec[3 * B:]

    def forward(self, data):
        mixed, noise, clean, full_noise = self.spectrograms(data)
        pred_noise, mask = self.net(mixed, noise)
        rec = transform.batch_fast_icRM_sigmoid(mixed, mask)
        self.last_rec = rec.detach()                           # (detached: a kept graph would pin last step's AccumulateGrad nodes)
        l1 = self.criterion(pred_noise, full_noise)
        l2 = self.criterion(rec, clean)
        return (pred_noise, mask), {"stage1": l1, "stage2": l2}

    def evaluate(self, dataloader):
        self.net.eval()
        tot, n = 0.0, 0
        with torch.no_grad():
            for data in dataloader:
                _, losses = self.forward(data)
                tot += float(losses["stage2"])
                n += 1
        return tot / max(n, 1)


class SIDAgent(BaseAgent):
    """Stage-1 (AudioVisualNet) trainer, M1/agent.py:144-237: BCEWithLogitsLoss on (B, 60) labels."""

    def build_net(self, config):
        return get_network().to(self.device)

    def set_loss_function(self):
        self.criterion = L.BCEWithLogitsLoss.apply

    def forward(self, data):
        d = self.device
        audio = _dev(data["audio"], d).float() if "audio" in data else transform.stft_batch(_dev(data["mixed_wave"], d))
        label = _dev(data["label"], d).float()
        logits = self.net(audio, label.shape[1])
        return logits, {"bce": self.criterion(logits, label)}

    def evaluate(self, dataloader):
        """frame accuracy at the 0.5 threshold (M1/agent.py:209-230)."""
        self.net.eval()
        ok, n = 0, 0
        with torch.no_grad():
            for data in dataloader:
                logits, _ = self.forward(data)
                pred = (torch.sigmoid(logits) >= 0.5).float()
                ok += int((pred == _dev(data["label"], self.device)).sum())
                n += pred.numel()
        return ok / max(n, 1)


class GraphedTrainStep(object):
    """BASELINE configs[1] -- one training step of BOTH models on a batch of waveforms -- as CUDA-graph replays.

        step = GraphedTrainStep(sid_agent, joint_agent, batch, length, sr, fps)
        out = step(mixed, clean, full_noise, bits, label)     # (B, L) fp32 x 3, (B, n_bits) uint8, (B, n_bits) fp32 device tensors
        out["losses"] (3,) = bce, stage1, stage2;  out["wave"] (B, 158 (T-1)) = iSTFT of the recovered spectrogram

    The eager step enqueues ~1000 kernels from Python (tens of ms of host time: the GPU would wait for the host on a slower CPU,
    and data-parallel ranks would wait for the slowest host).  Here the work is captured once, after `warmup` eager steps, into
    two graphs and replayed with one launch each:
        graph 1: silent-interval gate -> 4 x STFT (one launch) -> SID forward, BCE, backward (M1/agent.py:185-206)
        graph 2: JointModel forward, cRM recovery, 2 x MSE, backward (M2/agent.py:176-190, 101-105) -> iSTFT of the recovery
    The data-parallel exchange stays outside the graphs: SID's flat gradient is all-reduced (NCCL, its own stream) WHILE graph 2
    runs; the Joint gradient follows; then the two fused Adam updates (device-resident step clock, FlatAdam.apply).
    The returned tensors are static buffers of the graphs: consume (or copy) them before the next call."""

    def __init__(self, sid_agent, joint_agent, batch, length, sr=16000, fps=30.0, n_bits=None, warmup=3, concurrent=None):
        self.sid, self.joint = sid_agent, joint_agent
        # concurrent (SOS_CONCURRENT=1; default off): ONE graph in which the detector's step and the joint model's step are parallel
        # branches (they share only the spectrograms), so that one's kernels could fill the SMs the other leaves idle (LSTM recurrences
        # on 26 of 148 SMs, kernel tails).  Measured at batch 32: 79.6 ms against 79.5 ms for the two graphs in sequence -- the big
        # kernels are persistent whole-GPU grids, two of them only time-slice -- and the sequential form overlaps the detector's
        # gradient all-reduce with graph 2, so that one stays the default.
        self.concurrent = (os.environ.get("SOS_CONCURRENT", "0") == "1") if concurrent is None else bool(concurrent)
        dev = sid_agent.device
        self.ratio = sr / fps
        n_bits = int(round(length / sr * fps)) if n_bits is None else n_bits
        self.inp = {"mixed": torch.zeros(batch, length, device=dev), "clean": torch.zeros(batch, length, device=dev),
                    "full_noise": torch.zeros(batch, length, device=dev), "bits": torch.ones(batch, n_bits, device=dev, dtype=torch.uint8),
                    "label": torch.zeros(batch, n_bits, device=dev)}
        self.batch, self.warmup, self.calls = batch, warmup, 0
        self.g1 = self.g2 = None
        self.out = None
        self.launches_per_step = 0
        self.stream, self.stream2 = torch.cuda.Stream(), torch.cuda.Stream()
        self.world = torch.distributed.get_world_size() if (torch.distributed.is_available() and torch.distributed.is_initialized()) else 1

    def _part0(self):
        d = self.inp
        L.pack_all()                                             # (the first node of the step's graph: both models' weight operands)
        gated = tools.gate_noise(d["mixed"], self.ratio, d["bits"])
        self.spec = transform.stft_batch(torch.cat([d["mixed"], gated, d["clean"], d["full_noise"]]))

    def _sid_part(self):
        self.sid.net.train()
        _, l_sid = self.sid.forward({"audio": self.spec[:self.batch], "label": self.inp["label"]})
        self.sid.backward(l_sid)
        return l_sid["bce"].detach()

    def _part1(self):
        self._part0()
        return self._sid_part()

    def _part2(self):
        B, spec = self.batch, self.spec
        self.joint.net.train()
        _, l_jt = self.joint.forward({"mixed": spec[:B], "noise": spec[B:2 * B], "clean": spec[2 * B:3 * B], "full_noise": spec[3 * B:]})
        self.joint.backward(l_jt)
        wave = transform.istft_batch(self.joint.last_rec.detach())
        return l_jt["stage1"].detach(), l_jt["stage2"].detach(), wave

    def _whole(self):
        """The step with the detector's part as a parallel branch (on self.stream2) of the current stream."""
        cur = torch.cuda.current_stream()
        self._part0()
        self.stream2.wait_stream(cur)                            # fork: the detector's branch starts once the spectrograms exist
        with torch.cuda.stream(self.stream2):
            bce = self._sid_part()
        l1, l2, wave = self._part2()
        cur.wait_stream(self.stream2)                            # join
        return self._finish(bce, l1, l2, wave)

    def _finish(self, bce, l1, l2, wave):
        return {"losses": torch.stack([bce, l1, l2]), "wave": wave}

    def _exchange_and_update(self, h_sid=None):
        sid_o, jt_o = self.sid.optimizer, self.joint.optimizer
        if self.world > 1:
            h_jt = torch.distributed.all_reduce(jt_o.flat_grad, async_op=True)
            if h_sid is not None:
                h_sid.wait()
            h_jt.wait()
        sid_o.sync_clock()
        jt_o.sync_clock()
        sid_o.apply(self.world)
        jt_o.apply(self.world)
        L.weights_changed()
        for a in (self.sid, self.joint):
            a.clock.tick()

    def __call__(self, mixed, clean, full_noise, bits, label):
        for k, v in (("mixed", mixed), ("clean", clean), ("full_noise", full_noise), ("bits", bits), ("label", label)):
            self.inp[k].copy_(v, non_blocking=True)
        self.calls += 1
        if self.calls <= self.warmup:
            # eager warm-up (allocator pools, plan cache, lazily built tables) on a side stream: autograd's AccumulateGrad nodes must
            # not be born on the legacy default stream, or the capture below could not include them
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                if self.concurrent:
                    out = self._whole()                         # (same streams as the capture: its side streams / workspaces exist by then)
                    self._exchange_and_update(None)
                else:
                    bce = self._part1()
                    h = torch.distributed.all_reduce(self.sid.optimizer.flat_grad, async_op=True) if self.world > 1 else None
                    l1, l2, wave = self._part2()
                    out = self._finish(bce, l1, l2, wave)
                    self._exchange_and_update(h)
            cur.wait_stream(self.stream)
            return out
        if self.g1 is None:
            torch.cuda.synchronize()
            pool = torch.cuda.graph_pool_handle()
            from . import _lib
            n0 = _lib.launch_count
            # (thread_local: NCCL's watchdog thread polls CUDA events while we capture)
            if self.concurrent:
                self.g1, self.g2 = torch.cuda.CUDAGraph(), None
                with torch.cuda.graph(self.g1, pool=pool, stream=self.stream, capture_error_mode="thread_local"):
                    self.out = self._whole()
            else:
                self.g1, self.g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.g1, pool=pool, stream=self.stream, capture_error_mode="thread_local"):
                    bce = self._part1()
                with torch.cuda.graph(self.g2, pool=pool, stream=self.stream, capture_error_mode="thread_local"):
                    l1, l2, wave = self._part2()
                    self.out = self._finish(bce, l1, l2, wave)
            self.launches_per_step = _lib.launch_count - n0    # this library's kernels inside the two graphs
            for a in (self.sid, self.joint):                    # autograd replaced .grad views? re-attach (host only)
                a.optimizer.zero_grad_views()
        self.g1.replay()
        h = torch.distributed.all_reduce(self.sid.optimizer.flat_grad, async_op=True) if self.world > 1 else None
        if self.g2 is not None:
            self.g2.replay()
        self._exchange_and_update(h)
        return self.out


def get_agent(config):
    return SIDAgent(config) if getattr(config, "model", "joint") == "sid" else MyAgent(config)
