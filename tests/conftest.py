import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import sos_b200  # noqa: F401
    from sos_b200 import ops
    ops.init()
    return torch.device("cuda:0")


def record(name, **values):
    """Parity numbers of the GPU tests, captured (not only printed): one JSON line per call appended to
    gpurun_out/parity.jsonl (copied to profiles/parity_rNN.txt when a round's numbers are committed)."""
    import json
    out = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity.jsonl"), "a") as f:
            f.write(json.dumps({"test": name, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in values.items()}}) + "\n")
    except OSError:
        pass
