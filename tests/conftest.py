import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_rx_golden():
    meta = json.load(open(os.path.join(GOLDEN, "rx_cases.json")))
    arrs = np.load(os.path.join(GOLDEN, "rx_cases.npz"))
    return [(m, arrs[m["name"]]) for m in meta]


def load_tx_golden():
    meta = json.load(open(os.path.join(GOLDEN, "tx_cases.json")))
    arrs = np.load(os.path.join(GOLDEN, "tx_cases.npz"))
    return [(m, arrs[m["name"]] if m["name"] in arrs.files else None) for m in meta]


def load_gate_golden():
    return json.load(open(os.path.join(GOLDEN, "gate_cases.json")))


def load_gate_multi_golden():
    return json.load(open(os.path.join(GOLDEN, "gate_multi_cases.json")))


def gate_multi_stream(g, tx_frames):
    """Rebuilds the recorded stream of a gate_multi golden case from its recipe; ``tx_frames(msg,
    baud, training_time)`` must be sample-exact with the reference's Transmitter.save (checked by
    the case's SHA-256)."""
    import hashlib

    import numpy as np
    parts = []
    for lead, msg, gain, tt in g["segments"]:
        parts.append(np.zeros(lead, np.int16))
        if msg is not None:
            parts.append((tx_frames(msg.encode("utf-8"), g["baud"], tt).astype(np.float64) * gain).astype(np.int16))
    s = np.concatenate(parts)
    if g["cut"]:
        s = s[:len(s) - g["cut"]]
    assert len(s) == g["n"] and hashlib.sha256(s.astype("<i2").tobytes()).hexdigest() == g["sha256"], g["name"]
    return s


@pytest.fixture(scope="session")
def rx_golden():
    return load_rx_golden()


@pytest.fixture(scope="session")
def tx_golden():
    return load_tx_golden()


def has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False
