"""Golden stdout of the reference's ``Eval.Print_Every_class_Eval`` (utils/eval.py:90-106): the method is run
UNMODIFIED on the frozen confusion matrices of tests/golden/eval_tensors.npz and what it prints is stored verbatim in
tests/golden/eval_print.json.  Test infrastructure; needs /root/reference (build container only).

    python oracle/make_golden_print.py
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import make_golden  # noqa: E402  (load_reference: the reference's modules by path)


def main():
    _, ref_eval = make_golden.load_reference()
    t = np.load(os.path.join(ROOT, "tests", "golden", "eval_tensors.npz"))
    with open(os.path.join(ROOT, "tests", "golden", "eval_kats.json")) as f:
        cases = json.load(f)["cases"]
    out = []
    for c in cases:
        name, C = c["name"], c["C"]
        for flag in ([False, True] if C == 19 else [False]):
            ev = ref_eval.Eval(C)
            ev.confusion_matrix = t[name + "_cm"].astype(np.float64)
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf), np.errstate(all="ignore"):
                ev.Print_Every_class_Eval(out_16_13=flag)
            out.append(dict(name=name, C=C, out_16_13=flag, stdout=buf.getvalue()))
    with open(os.path.join(ROOT, "tests", "golden", "eval_print.json"), "w") as f:
        json.dump(dict(numpy=np.__version__, cases=out), f, indent=1)
    print(f"{len(out)} printouts frozen")


if __name__ == "__main__":
    main()
