"""agent.GraphedTrainStep (the whole training step of both models as two CUDA-graph replays, device-resident Adam clock) must do
what the eager agents do: same losses step by step on changing data, the same parameters afterwards, a working learning-rate
change, and a checkpoint step count that follows the replays."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


LR = 1e-5          # small steps: Adam turns last-bit gradient noise (fp32 atomics order) into up to 2 lr per weight and step, and at the
                   # reference's 1e-3 two independently updated agent pairs drift apart chaotically within a few steps


def _agents():
    from sos_b200 import agent as ag
    torch.manual_seed(0)
    sid = ag.SIDAgent(ag.default_config(model="sid", lr=LR))
    torch.manual_seed(1)
    joint = ag.MyAgent(ag.default_config(model="joint", lr=LR))
    return sid, joint


@pytest.mark.parametrize("concurrent", [False, True])
def test_graphed_step_matches_eager(cuda, concurrent):
    """(concurrent = False: two graphs in sequence, the default.  True: one graph with the detector's step as a parallel branch --
    an opt-in mode (SOS_CONCURRENT=1, no measured gain): tested only with SOS_TEST_CONCURRENT=1.)
    Three identically initialised agent pairs see the same batches: two run eagerly (agent.train_func), one through
    GraphedTrainStep (2 eager warm-up steps, the capture, then replays).  The weight-gradient kernels merge their pixel slices with
    fp32 atomics, so even the two EAGER pairs drift apart from step to step; the graphed pair may differ from an eager pair by at
    most 3 x what the eager pairs differ from each other (plus a floor), in the three losses and in the recovered waveform."""
    import os
    if concurrent and os.environ.get("SOS_TEST_CONCURRENT") != "1":
        pytest.skip("opt-in mode: set SOS_TEST_CONCURRENT=1")
    from sos_b200 import agent as ag, tools, transform, networks
    from oracle import synth
    B, L, STEPS = 2, 16000, 6
    ratio = 16000 / 30.0
    pairs = [_agents(), _agents()]
    sid_g, joint_g = _agents()
    step = ag.GraphedTrainStep(sid_g, joint_g, B, L, 16000, 30.0, warmup=2, concurrent=concurrent)
    worst = [0.0, 0.0, 0.0, 0.0]
    for i in range(STEPS):
        clips = synth.make_batch(B, length=L, start=10 * i)
        w = {k: torch.tensor(clips[k], device=cuda) for k in ("mixed", "clean", "full_noise")}
        bits = torch.tensor(np.array([[int(c) for c in b] for b in clips["bits"]], dtype=np.uint8), device=cuda)
        lab = torch.tensor(clips["label"], device=cuda)
        if i == 4:                                        # a learning-rate change between replays (StepLR) must reach the device clock
            for a in (sid_g, joint_g, *pairs[0], *pairs[1]):
                a.optimizer.param_groups[0]["lr"] = LR / 4
        out = step(w["mixed"], w["clean"], w["full_noise"], bits, lab)
        got = out["losses"].cpu().numpy().astype(np.float64)
        wave_g = out["wave"].clone()
        gated = tools.gate_noise(w["mixed"], ratio, bits)
        spec = transform.stft_batch(torch.cat([w["mixed"], gated, w["clean"], w["full_noise"]]))
        ref = []
        for sid_e, joint_e in pairs:
            _, ls = sid_e.train_func({"audio": spec[:B], "label": lab})
            _, lj = joint_e.train_func({"mixed": spec[:B], "noise": spec[B:2 * B], "clean": spec[2 * B:3 * B], "full_noise": spec[3 * B:]})
            ref.append((np.array([float(ls["bce"].detach()), float(lj["stage1"].detach()), float(lj["stage2"].detach())]),
                        transform.istft_batch(joint_e.last_rec)))
        d_ee = float(np.abs(ref[0][0] - ref[1][0]).max() / np.abs(ref[0][0]).max())
        d_ge = float(np.abs(got - ref[0][0]).max() / np.abs(ref[0][0]).max())
        w_ee = float((ref[0][1] - ref[1][1]).abs().mean())
        w_ge = float((wave_g - ref[0][1]).abs().mean())
        worst = [max(a, b) for a, b in zip(worst, (d_ee, d_ge, w_ee, w_ge))]
        print(f"step {i}: loss drift eager/eager {d_ee:.2e} graph/eager {d_ge:.2e}; wave drift eager/eager {w_ee:.2e} graph/eager {w_ge:.2e}")
        # (the eager/eager drift itself scatters between 2e-4 and 1e-2 from step to step and run to run -- atomics order amplified by
        # BatchNorm over 2 clips -- so the bound uses the largest drift seen so far and a floor inside that scatter)
        assert d_ge <= 3 * worst[0] + 6e-3, (i, got, ref[0][0], ref[1][0])
        # (the first Adam step moves every weight by lr * sign(g): where the weight-gradient atomics decide a sign, or a gradient's
        #  power-of-two half scale, an agent pair lands in one of a few discrete states -- 1.08e-4 of wave drift apart at step 1,
        #  measured over sixteen runs: sometimes the two eager pairs differ by it, sometimes only the graphed pair does, and pairs in
        #  different states then drift apart by 7e-4 / 1.5e-3 / 1.1e-3 / 3e-3 over the next steps while pairs in the same state
        #  stay 3-5 x closer.  The floor per step is ~3 x the different-state drift; the bound also uses the largest eager/eager
        #  drift seen so far)
        assert w_ge <= 3 * worst[2] + (1e-5, 5e-4, 2e-3, 4e-3, 4e-3, 8e-3)[i], (i, w_ge, w_ee, worst)
    assert step.g1 is not None and (step.g2 is None) == concurrent and step.launches_per_step > 100
    assert sid_g.optimizer.step_count == STEPS and abs(float(sid_g.optimizer.state[1]) - STEPS) < 1e-6
    assert abs(float(joint_g.optimizer.state[0]) - LR / 4) < 1e-12
    for (k, p), (_, q) in zip(pairs[0][1].net.state_dict().items(), joint_g.net.state_dict().items()):
        if not p.is_floating_point():
            assert torch.equal(p, q), k
    torch.manual_seed(1)
    fresh = networks.get_network(object()).state_dict()["stage2.fc.4.weight"].to(cuda)
    moved = float((joint_g.net.state_dict()["stage2.fc.4.weight"] - fresh).abs().mean())
    moved_e = float((pairs[0][1].net.state_dict()["stage2.fc.4.weight"] - fresh).abs().mean())
    assert 0.5 * moved_e < moved < 2 * moved_e and moved > LR, (moved, moved_e)         # the replays DID update the parameters like the eager steps
