"""The evaluation port (oracle/eval_port.py) against vectors frozen from the reference's OWN per-image loop
(tools/analysis.py:171-240, resultEvaluater.getvalResult, executed by oracle/make_golden_perimage.py)."""
import json
import os

import numpy as np
import pytest

from oracle import eval_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def perimage_cases():
    with open(os.path.join(GOLDEN, "perimage_kats.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("case", perimage_cases(), ids=lambda c: c["name"])
def test_port_reproduces_the_per_image_loop(case):
    t = np.load(os.path.join(GOLDEN, "perimage_tensors.npz"))
    name, C = case["name"], case["C"]
    total = eval_port.EvalPort(C)
    for b, row in enumerate(case["per_image"]):
        label = t[f"{name}/label{b}"].astype(np.int64)
        arg = np.argmax(t[f"{name}/pred{b}"], axis=1)                 # tools/analysis.py:207
        one = eval_port.EvalPort(C)
        one.add_batch(label, arg)
        total.add_batch(label, arg)
        assert one.Pixel_Accuracy() == row["PA"]
        assert list(one.Mean_Pixel_Accuracy()) == row["MPA"]
        assert one.Mean_Intersection_over_Union()[0] == row["MIoU"]
        assert list(one.Frequency_Weighted_Intersection_over_Union()) == row["FWIoU"]
    assert np.array_equal(total.confusion_matrix, t[f"{name}/total_cm"])
    assert list(total.Mean_Intersection_over_Union()) == case["total_miou"]
