"""The oracle's source-side step (oracle/loss_port.chain_source, oracle/loss_math.source_ce,
oracle/eval_port) against the vectors frozen from torch's CrossEntropyLoss + the reference's own Eval
(tests/golden/source_*, produced by oracle/make_golden_source.py).  CPU only."""
import hashlib
import json
import math
import os

import numpy as np
import pytest
import torch

from maxsquareloss_b200 import synth
from oracle import eval_port, loss_math, loss_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "source_kats.json")) as _f:
    SOURCE = json.load(_f)["cases"]


def inputs(c):
    lo, y = synth.source_case(c["N"], c["C"], c["hw"], c["HW"], c["seed"], c["scale"], c["label_kind"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"]
    assert hashlib.sha256(y.numpy().tobytes()).hexdigest() == c["label_sha256"]
    return lo, y


def close(a, b, rtol):
    return math.isnan(a) if math.isnan(b) else abs(a - b) <= rtol * abs(b)


@pytest.mark.parametrize("c", SOURCE, ids=[c["name"] for c in SOURCE])
def test_port_and_closed_form_reproduce_reference(c):
    lo, y = inputs(c)
    r = loss_port.chain_source(lo, y, c["C"], c["grad_scale"])
    assert close(r["loss"].item(), c["loss"], 1e-6) and r["nvalid"] == c["nvalid"]
    assert hashlib.sha256(r["argpred"].astype(np.int64).tobytes()).hexdigest() == c["argpred_sha256"]
    ev = eval_port.EvalPort(c["C"])
    ev.add_batch(y.numpy(), r["argpred"])
    assert hashlib.sha256(ev.confusion_matrix.astype(np.int64).tobytes()).hexdigest() == c["cm_sha256"]
    m = loss_math.source_ce(lo.numpy(), y.numpy(), c["grad_scale"])
    assert hashlib.sha256(m["argpred"].astype(np.int64).tobytes()).hexdigest() == c["argpred_sha256"]
    assert close(float(m["loss"]), c["loss"], 1e-5) and m["nvalid"] == c["nvalid"]
    g = m["grad_logits"]
    if c["nvalid"]:
        assert close(float(np.abs(g).sum()), c["grad_sum_abs"], 1e-4)
        assert close(float(np.sqrt((g * g).sum())), c["grad_l2"], 1e-4)
        assert close(r["grad"].abs().sum().item(), c["grad_sum_abs"], 1e-5)
    else:
        assert not g.any() and not r["grad"].any()
