"""The MinEnt restatement (oracle/loss_port.py: soft_cross_entropy / iw_soft_cross_entropy) against vectors frozen from the
reference's own classes run on targets that are NOT softmax(inputs) (utils/loss.py:17-67, oracle/make_golden_softce.py)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import loss_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def softce_cases():
    with open(os.path.join(GOLDEN, "softce_kats.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("c", softce_cases(), ids=lambda c: c["name"])
def test_port_reproduces_the_reference_classes(c):
    t = np.load(os.path.join(GOLDEN, "softce_tensors.npz"))
    z = torch.from_numpy(t[c["name"] + "__inputs"])
    assert hashlib.sha256(z.numpy().tobytes()).hexdigest() == c["input_sha256"]
    x = z.clone().requires_grad_(True)
    if c["target"] == "self":
        tt = F.softmax(x, dim=1)
        tt.retain_grad()
    else:
        tt = torch.from_numpy(t[c["name"] + "__target"]).clone().requires_grad_(True)
    if c["iw"]:
        loss, hist = loss_port.iw_soft_cross_entropy(x, tt, c["C"], c["ratio"], return_aux=True)
        assert hist.tolist() == c["hist"]
    else:
        loss = loss_port.soft_cross_entropy(x, tt)
    (c["grad_scale"] * loss).backward()
    assert loss.item() == c["loss"]                                   # same ops, same library: identical bits
    assert np.array_equal(x.grad.numpy(), t[c["name"] + "__grad_inputs"])
    assert np.array_equal(tt.grad.numpy(), t[c["name"] + "__grad_target"])
