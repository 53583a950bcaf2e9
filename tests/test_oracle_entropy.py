"""The oracle's MinEnt losses (oracle/loss_port.chain_entropy, oracle/loss_math.fused_entropy) against the
vectors frozen from the reference's softCrossEntropy / IWsoftCrossEntropy (tests/golden/entropy_*, produced
by oracle/make_golden_entropy.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

from maxsquareloss_b200 import synth
from oracle import loss_math, loss_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "entropy_kats.json")) as _f:
    ENT = json.load(_f)["cases"]


def logits(c):
    lo = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"], c["quantize"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"]
    return lo


@pytest.mark.parametrize("c", ENT, ids=[c["name"] for c in ENT])
def test_port_and_closed_form_reproduce_reference(c):
    lo = logits(c)
    loss, grad, hist = loss_port.chain_entropy(lo, c["HW"], c["C"], c["iw"], c["ratio"], c["grad_scale"])
    assert abs(loss.item() - c["loss"]) <= 1e-6 * abs(c["loss"])
    assert abs(grad.abs().sum().item() - c["grad_sum_abs"]) <= 1e-5 * c["grad_sum_abs"]
    if c["iw"]:
        assert hist.tolist() == c["hist"]
    r = loss_math.fused_entropy(lo.numpy(), c["HW"], c["C"], c["iw"], c["ratio"], c["grad_scale"])
    assert abs(r["loss"] - c["loss"]) <= 1e-5 * abs(c["loss"])
    g = r["grad_logits"]
    assert abs(np.abs(g).sum() - c["grad_sum_abs"]) <= 1e-4 * c["grad_sum_abs"]
    assert abs(np.sqrt((g * g).sum()) - c["grad_l2"]) <= 1e-4 * c["grad_l2"]
    if c["iw"]:
        assert r["hist"].tolist() == c["hist"]            # bit-exact, incl. the case with exact ties


def test_closed_form_gradient_elementwise():
    t = np.load(os.path.join(GOLDEN, "entropy_tensors.npz"))
    n = 0
    for c in ENT:
        if c["name"] + "__grad" not in t.files:
            continue
        r = loss_math.fused_entropy(t[c["name"] + "__logits"], c["HW"], c["C"], c["iw"], c["ratio"], c["grad_scale"])
        ref = t[c["name"] + "__grad"].astype(np.float64)
        assert np.abs(r["grad_logits"] - ref).max() <= 1e-4 * np.abs(ref).max()
        n += 1
    assert n >= 2
