"""CPU, world_size 2 over gloo: the host logic of the data-parallel step.  Each rank owns its clips end to
end; the only exchange is ONE all-reduce of the flat gradient buffer, whose sum / world equals the gradient of the full-batch
mean loss (what the reference's nn.DataParallel computes, M2/agent.py:159-161) when the shards are equal."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(5)
    return torch.nn.Sequential(torch.nn.Linear(6, 7), torch.nn.Tanh(), torch.nn.Linear(7, 3))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sos_b200  # noqa: F401
    from sos_b200.agent import FlatAdam
    import bench
    net = _model()
    opt = FlatAdam(net.parameters(), 1e-3)
    g = torch.Generator().manual_seed(9)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    per = 8 // world
    xs, ys = x[rank * per:(rank + 1) * per], y[rank * per:(rank + 1) * per]
    opt.zero_grad()
    torch.nn.functional.mse_loss(net(xs), ys).backward()
    # autograd accumulates into the views of the flat buffer
    assert all(p.grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * o for p, o in zip(opt.params, opt.offsets))
    w = opt.exchange()
    assert w == world
    np.save(os.path.join(out_dir, f"grad{rank}.npy"), (opt.flat_grad / w).numpy())
    # shard assignment of the benchmark: rank r synthesises clips [r*B, (r+1)*B)
    clips = bench.synth_batch(2, length=4000, start=rank * 2)
    np.save(os.path.join(out_dir, f"clips{rank}.npy"), clips["mixed"])
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_equals_full_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    net = _model()
    g = torch.Generator().manual_seed(9)
    x, y = torch.randn(8, 6, generator=g), torch.randn(8, 3, generator=g)
    torch.nn.functional.mse_loss(net(x), y).backward()
    want = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 4)) for p in net.parameters()]).numpy()
    g0, g1 = np.load(tmp_path / "grad0.npy"), np.load(tmp_path / "grad1.npy")
    assert np.array_equal(g0, g1), "ranks must hold identical reduced gradients (weights stay bit-identical without a broadcast)"
    assert np.abs(g0 - want).max() < 1e-6
    c0, c1 = np.load(tmp_path / "clips0.npy"), np.load(tmp_path / "clips1.npy")
    import bench
    whole = bench.synth_batch(4, length=4000)["mixed"]
    assert np.array_equal(np.concatenate([c0, c1]), whole), "the ranks' shards tile the global batch"


def test_flat_adam_checkpoint_roundtrip():
    """optimizer_state_dict keeps torch.optim.Adam's layout (M2/agent.py:65-77 saves it verbatim)."""
    import sos_b200  # noqa: F401
    from sos_b200.agent import FlatAdam, StepLR
    net = _model()
    opt = FlatAdam(net.parameters(), 1e-3)
    opt.step_count = 3
    opt.exp_avg.normal_()
    opt.exp_avg_sq.uniform_()
    sd = opt.state_dict()
    ref = torch.optim.Adam(_model().parameters(), 1e-3)
    assert set(sd["param_groups"][0]) >= {"lr", "betas", "eps", "params"} and len(sd["state"]) == len(list(net.parameters()))
    ref.load_state_dict({"state": {k: dict(v) for k, v in sd["state"].items()}, "param_groups": [dict(sd["param_groups"][0], maximize=False, foreach=None,
                        capturable=False, differentiable=False, fused=None)]})          # torch's own Adam accepts it
    opt2 = FlatAdam(_model().parameters(), 5e-4)
    opt2.load_state_dict(sd)
    sd2 = opt2.state_dict()
    assert opt2.step_count == 3 and opt2.param_groups[0]["lr"] == 1e-3
    assert all(torch.equal(sd2["state"][i]["exp_avg"], sd["state"][i]["exp_avg"]) and
               torch.equal(sd2["state"][i]["exp_avg_sq"], sd["state"][i]["exp_avg_sq"]) for i in sd["state"])
    sch = StepLR(opt2, 15)
    for e in range(1, 16):
        sch.step(e)
    assert abs(opt2.param_groups[0]["lr"] - 1e-4) < 1e-12
