"""End-to-end GPU parity of the two networks against the golden vectors generated from the reference itself
(oracle/make_golden.py): logits / n_pred / mask / losses / selected gradients, eval and train mode."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

MASK_L1_TOL = 1e-3            # north_star: mask L1 vs reference <= 1e-3


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(golden_dir + "/nets.npz")


def _gradcheck(params, gold, prefix, rtol):
    for key in gold.files:
        if not key.startswith(prefix + "_grad:"):
            continue
        name = key.split(":", 1)[1]
        want = gold[key]
        got = params[name].grad.cpu().numpy()
        if got.shape != want.shape:
            got = got[:16, :16]
        scale = np.abs(want).max() + 1e-12
        err = np.abs(got - want).max() / scale
        print(f"  grad {name}: rel err {err:.2e}")
        assert err < rtol, (name, err)


@pytest.mark.parametrize("mode", ["eval", "train"])
def test_sid_matches_reference(cuda, gold, mode):
    from sos_b200 import networks, layers as L
    from oracle import nets
    sid = networks.get_network()
    sid.load_state_dict(nets.synth_state_dict(nets.sid_shapes(), 3))
    sid = sid.to(cuda).train(mode == "train")
    x, lab = torch.tensor(gold["x"], device=cuda), torch.tensor(gold["label"], device=cuda)
    logits = sid(x, lab.shape[1])
    loss = L.BCEWithLogitsLoss.apply(logits, lab)
    loss.backward()
    want = gold[f"sid_{mode}_logits"]
    err = np.abs(logits.detach().cpu().numpy() - want).max()
    print(f"sid {mode}: logits max err {err:.2e} (scale {np.abs(want).max():.2f}) loss {float(loss):.6f} vs {float(gold[f'sid_{mode}_loss']):.6f}")
    assert err < 2e-2 * max(1.0, np.abs(want).max())
    assert abs(float(loss) - float(gold[f"sid_{mode}_loss"])) < 2e-3
    _gradcheck(dict(sid.named_parameters()), gold, f"sid_{mode}", 5e-2)
    if mode == "train":
        rm = sid.state_dict()["encoder_audio.3.block.1.running_mean"].cpu().numpy()
        assert np.abs(rm - gold["sid_train_rm:encoder_audio.3.block.1.running_mean"]).max() < 1e-3
    if mode == "eval":
        with torch.no_grad():
            fast = sid(x, lab.shape[1])                        # fused-epilogue inference path
        assert float((fast - logits.detach()).abs().max()) < 2e-2 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("tag", ["plain", "spread"])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_joint_matches_reference(cuda, gold, mode, tag):
    from sos_b200 import networks, layers as L, transform
    from oracle import nets
    joint = networks.get_network(object())
    joint.load_state_dict(nets.synth_state_dict(nets.joint_shapes(), 4, spread=(tag == "spread")))
    joint = joint.to(cuda).train(mode == "train")
    x, n = torch.tensor(gold["x"], device=cuda), torch.tensor(gold["n"], device=cuda)
    n_pred, mask = joint(x, n)
    rec = transform.batch_fast_icRM_sigmoid(x, mask)
    l1 = L.MSELoss.apply(n_pred, torch.tensor(gold["tgt_n"], device=cuda))
    l2 = L.MSELoss.apply(rec, torch.tensor(gold["tgt_c"], device=cuda))
    (l1 + l2).backward()
    pre = f"joint_{tag}_{mode}"
    e_np = np.abs(n_pred.detach().cpu().numpy() - gold[pre + "_npred"]).mean()
    e_mask = np.abs(mask.detach().cpu().numpy() - gold[pre + "_mask"]).mean()
    spread = gold[pre + "_mask"].max() - gold[pre + "_mask"].min()
    print(f"{pre}: n_pred L1 {e_np:.2e}  mask L1 {e_mask:.2e} (mask range {spread:.3f})  loss1 {float(l1):.5f}/{float(gold[pre + '_loss1']):.5f}"
          f" loss2 {float(l2):.4f}/{float(gold[pre + '_loss2']):.4f}")
    assert e_mask < MASK_L1_TOL
    assert e_np < 1e-2 * np.abs(gold[pre + "_npred"]).mean() + 1e-4
    assert abs(float(l1) - float(gold[pre + "_loss1"])) < 1e-2 * float(gold[pre + "_loss1"])
    if tag == "plain":
        assert abs(float(l2) - float(gold[pre + "_loss2"])) < 2e-2 * float(gold[pre + "_loss2"])
        _gradcheck(dict(joint.named_parameters()), gold, pre, 8e-2)
    if mode == "eval":
        with torch.no_grad():
            n2, m2 = joint(x, n)
        assert float((m2 - mask.detach()).abs().mean()) < MASK_L1_TOL
