"""``Print_Every_class_Eval`` (utils/eval.py:90-106) of the product's metric code against the
reference's own stdout, frozen by oracle/make_golden_print.py.  Host-only: the metrics are float64 NumPy on the matrix."""
import json
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "eval_print.json")) as _f:
    CASES = json.load(_f)["cases"]


@pytest.mark.parametrize("case", CASES, ids=[f"{c['name']}-{int(c['out_16_13'])}" for c in CASES])
def test_print_every_class_matches_the_reference_stdout(case, capsys):
    from maxsquareloss_b200.eval import HostEval
    cm = np.load(os.path.join(GOLDEN, "eval_tensors.npz"))[case["name"] + "_cm"]
    HostEval(case["C"], cm).Print_Every_class_Eval(out_16_13=case["out_16_13"])
    assert capsys.readouterr().out == case["stdout"]
