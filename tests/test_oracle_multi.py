"""The oracle's multi-level guidance (oracle/loss_port.chain_multi, oracle/loss_math.guidance)
against the vectors frozen from the reference's own train_target source
(tests/golden/multi_*, produced by oracle/make_golden_multi.py).  CPU only."""
import hashlib
import json
import math
import os

import numpy as np
import pytest
import torch

from maxsquareloss_b200 import synth
from oracle import loss_math, loss_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "multi_kats.json")) as _f:
    MULTI = json.load(_f)["cases"]


def heads(c):
    lo1 = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"])
    lo2 = synth.second_head(lo1, c["seed"])
    assert hashlib.sha256(lo1.numpy().tobytes()).hexdigest() == c["input1_sha256"]
    assert hashlib.sha256(lo2.numpy().tobytes()).hexdigest() == c["input2_sha256"]
    return lo1, lo2


def close(a, b, rtol):
    if math.isnan(b):
        return math.isnan(a)
    return abs(a - b) <= rtol * abs(b)


@pytest.mark.parametrize("c", MULTI, ids=[c["name"] for c in MULTI])
def test_port_reproduces_reference_train_target(c):
    lo1, lo2 = heads(c)
    r = loss_port.chain_multi(lo1, lo2, c["HW"], c["C"], c["kind"], c["ratio"], c["threshold"], c["lambda_target"],
                              c["lambda_seg"])
    assert hashlib.sha256(r["label_2"].numpy().tobytes()).hexdigest() == c["label2_sha256"]
    assert r["nvalid"] == c["nvalid"]
    assert close(r["loss_target"].item(), c["loss_target"], 1e-6)
    assert close(r["loss_target_2"].item(), c["loss_target_2"], 1e-6)
    assert close((r["loss_target"] + r["loss_target_2"]).item(), c["loss_total"], 1e-6)
    if c["nvalid"]:
        assert close(r["grad1"].abs().sum().item(), c["grad1_sum_abs"], 1e-5)
        assert close(r["grad2"].abs().sum().item(), c["grad2_sum_abs"], 1e-5)
        assert close(r["grad2"].norm().item(), c["grad2_l2"], 1e-5)
    else:
        assert not r["grad2"].any()        # CE over zero valid pixels: NaN loss but an all-zero gradient, as in torch


@pytest.mark.parametrize("c", MULTI, ids=[c["name"] for c in MULTI])
def test_closed_form_guidance_matches_reference(c):
    lo1, lo2 = heads(c)
    gs = c["lambda_seg"] * c["lambda_target"]
    r = loss_math.guidance(lo1.numpy(), lo2.numpy(), c["HW"], c["threshold"], gs)
    assert hashlib.sha256(r["label_2"].astype(np.int64).tobytes()).hexdigest() == c["label2_sha256"]
    assert r["nvalid"] == c["nvalid"]
    assert np.bincount(r["label_2"].reshape(-1) + 1, minlength=c["C"] + 1).tolist() == c["label2_hist"]
    assert close(float(gs * r["loss2"]), c["loss_target_2"], 1e-5)
    if c["nvalid"]:
        g = r["grad_logits2"]
        assert close(float(np.abs(g).sum()), c["grad2_sum_abs"], 1e-4)
        assert close(float(np.sqrt((g * g).sum())), c["grad2_l2"], 1e-4)


def test_closed_form_guidance_gradient_elementwise():
    t = np.load(os.path.join(GOLDEN, "multi_tensors.npz"))
    for c in MULTI:
        key = c["name"] + "__grad2"
        if key not in t.files or not c["nvalid"]:
            continue
        gs = c["lambda_seg"] * c["lambda_target"]
        r = loss_math.guidance(t[c["name"] + "__logits1"], t[c["name"] + "__logits2"], c["HW"], c["threshold"], gs)
        assert np.array_equal(r["label_2"], t[c["name"] + "__label2"].astype(np.int64))
        ref = t[key].astype(np.float64)
        assert np.abs(r["grad_logits2"] - ref).max() <= 1e-4 * np.abs(ref).max()
