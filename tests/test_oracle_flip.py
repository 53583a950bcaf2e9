"""The flip-ensemble restatement (oracle/eval_port.py:flip_ensemble_argmax) against vectors frozen from the reference's OWN
``Evaluater.validate`` run with ``--flip`` (tools/evaluate.py:98-202, executed unmodified by oracle/make_golden_flip.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import eval_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def flip_cases():
    with open(os.path.join(GOLDEN, "flip_kats.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", flip_cases(), ids=lambda c: c["name"])
def test_port_reproduces_the_reference_validate(case):
    t = np.load(os.path.join(GOLDEN, "flip_tensors.npz"))
    name, C = case["name"], case["C"]
    port = eval_port.EvalPort(C)
    for b in range(case["batches"]):
        arg = eval_port.flip_ensemble_argmax(torch.from_numpy(t[f"{name}/pred{b}"]), torch.from_numpy(t[f"{name}/pred_flip{b}"]))
        port.add_batch(t[f"{name}/label{b}"].astype(np.int64), arg)
    assert np.array_equal(port.confusion_matrix, t[f"{name}/cm"])
    assert int(port.confusion_matrix.sum()) == case["cm_sum"]
    miou = port.Mean_Intersection_over_Union()
    assert (list(miou) if isinstance(miou, tuple) else miou) == case["miou"]
    assert case["min_relative_gap"] > 2e-4            # the frozen argmax does not hinge on anyone's last ulp
