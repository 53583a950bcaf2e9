import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def loss_kats():
    with open(os.path.join(GOLDEN, "loss_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def eval_kats():
    with open(os.path.join(GOLDEN, "eval_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def loss_tensors():
    return np.load(os.path.join(GOLDEN, "loss_tensors.npz"))


@pytest.fixture(scope="session")
def eval_tensors():
    return np.load(os.path.join(GOLDEN, "eval_tensors.npz"))
