"""The oracle's ``--target_mode hard`` restatement (oracle/loss_port.chain_hard) against the vectors frozen from the
reference's own train_target source (tests/golden/hard_*, oracle/make_golden_hard.py).  CPU only."""
import hashlib
import json
import math
import os

import pytest

from maxsquareloss_b200 import synth
from oracle import loss_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "hard_kats.json")) as _f:
    HARD = json.load(_f)["cases"]


@pytest.mark.parametrize("c", HARD, ids=[c["name"] for c in HARD])
def test_port_reproduces_reference_hard_mode(c):
    lo = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"]
    r = loss_port.chain_hard(lo, c["HW"], c["threshold"], c["lambda_target"])
    assert hashlib.sha256(r["label"].numpy().tobytes()).hexdigest() == c["label_sha256"]
    assert r["nvalid"] == c["nvalid"]
    if math.isnan(c["loss_target"]):
        assert math.isnan(r["loss_target"].item()) and not r["grad"].any()
    else:
        assert abs(r["loss_target"].item() - c["loss_target"]) <= 1e-6 * abs(c["loss_target"])
        assert abs(r["grad"].abs().sum().item() - c["grad_sum_abs"]) <= 1e-5 * c["grad_sum_abs"]
        assert abs(r["grad"].norm().item() - c["grad_l2"]) <= 1e-5 * c["grad_l2"]
