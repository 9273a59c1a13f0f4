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
