"""Generate ``tests/golden/flip_*`` by executing the reference's OWN ``Evaluater.validate`` source with ``--flip``
(``/root/reference/tools/evaluate.py:98-202``: softmax of the prediction and of the prediction for the mirrored image, the
latter flipped back, averaged, ``np.argmax``, ``Eval.add_batch``).

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_flip``

``tools/evaluate.py`` cannot be imported here (``distutils``, ``tensorboardX``, dataset construction at import), so the
method's source text is cut out of the file with ``ast`` and compiled unmodified; it runs against a stub ``self`` carrying
exactly what the method reads: ``Eval`` = the reference's own ``utils/eval.py:Eval``, a two-batch ``dataloader``, a small
fixed convolutional ``model`` (NOT mirror-symmetric, so the flipped view really differs), ``args.flip = True``.
What is frozen: the logits the model returned for the image and for its mirror image, the labels, and the reference's
confusion matrix / metrics.  Cases are kept only if no pixel's two best averaged probabilities are within 2e-4 relative, so
that the frozen argmax does not depend on whose exponential is used (fp32 softmax implementations differ by ~1e-7 relative;
the CUDA kernel replays torch's arithmetic below 1e-5).
"""
import ast
import contextlib
import io
import json
import logging
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from tqdm import tqdm

from .make_golden import OUT, REF, load_reference, sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import synth  # noqa: E402


def reference_validate():
    path = os.path.join(REF, "tools", "evaluate.py")
    src = open(path).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "Evaluater")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "validate")
    lines = src.splitlines()[fn.lineno - 1:fn.end_lineno]
    ns = {"torch": torch, "F": F, "np": np, "tqdm": tqdm}
    exec(compile(textwrap.dedent("\n".join(lines)), path, "exec"), ns)
    return ns["validate"], (fn.lineno, fn.end_lineno)


class RecordingModel(nn.Module):
    """3x3 convolution with fixed random weights (asymmetric kernel): returns (pred, pred_2) like DeeplabMulti and
    remembers every output."""

    def __init__(self, C, seed, scale):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.conv = nn.Conv2d(3, C, 3, padding=1)
        with torch.no_grad():
            self.conv.weight.copy_(torch.randn(self.conv.weight.shape, generator=g) * scale)
            self.conv.bias.copy_(torch.randn(C, generator=g))
        self.outputs = []

    def forward(self, x):
        y = self.conv(x)
        self.outputs.append(y.detach().clone())
        return y, y * 0.5


FLIP_CASES = [
    # name, C, (N, H, W), seed, scale
    ("flip_c19_even", 19, (2, 24, 48), 41, 0.3),
    ("flip_c13_even", 13, (1, 24, 48), 42, 0.5),
    ("flip_c16_odd_width", 16, (1, 20, 33), 43, 0.4),
]


def min_relative_gap(pred, pred_flip):
    p = (F.softmax(pred, 1) + torch.flip(F.softmax(pred_flip, 1), dims=[-1])) / 2
    top = torch.topk(p, 2, dim=1).values
    return float(((top[:, 0] - top[:, 1]) / top[:, 0]).min())


def run_case(validate, ref_eval, case, keep):
    name, C, (N, H, W), seed, scale = case
    for attempt in range(400):                       # look for a seed without a near-tie pixel
        g = torch.Generator().manual_seed(seed + 1000 * attempt)
        batches = []
        for b in range(2):
            x = torch.randn(N, 3, H, W, generator=g)
            y = synth.blocky_labels(N, (H, W), C, seed + b, grid=(4, 8)).unsqueeze(1).float()    # datasets emit float labels (N,1,H,W)
            batches.append((x, y, [f"img{b}"]))
        model = RecordingModel(C, seed + 1000 * attempt, scale)
        stub = types.SimpleNamespace(
            logger=logging.getLogger("golden_flip"), Eval=ref_eval.Eval(C), current_epoch=0, cuda=False, device=torch.device("cpu"),
            dataloader=types.SimpleNamespace(val_loader=batches, valid_iterations=len(batches)), model=model,
            args=types.SimpleNamespace(flip=True, image_summary=False, class_16=(C == 16), multi=False))
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            metrics = validate(stub)
        preds = model.outputs[0::2]                  # model(x), model(flip(x)) per batch
        flips = model.outputs[1::2]
        gap = min(min_relative_gap(p, f) for p, f in zip(preds, flips))
        if gap > 2e-4:
            break
    else:
        raise RuntimeError(f"{name}: no seed without a near-tie pixel")
    cm = stub.Eval.confusion_matrix
    rec = {"name": name, "C": C, "shape": [N, H, W], "batches": len(batches), "attempt": attempt, "min_relative_gap": gap,
           "cm_sha": sha(cm.astype(np.int64)), "cm_sum": int(cm.sum()), "cm_diag": [int(v) for v in np.diag(cm)],
           "metrics": [float(m) for m in metrics],
           "miou": (lambda m: [float(v) for v in m] if isinstance(m, tuple) else float(m))(stub.Eval.Mean_Intersection_over_Union())}
    for b, ((x, y, _), p, f) in enumerate(zip(batches, preds, flips)):
        keep[f"{name}/pred{b}"] = p.numpy()
        keep[f"{name}/pred_flip{b}"] = f.numpy()
        keep[f"{name}/label{b}"] = y.squeeze(1).long().numpy().astype(np.int16)
    keep[f"{name}/cm"] = cm.astype(np.int64)
    return rec


def main():
    logging.getLogger("golden_flip").addHandler(logging.NullHandler())
    logging.getLogger("golden_flip").propagate = False
    _, ref_eval = load_reference()
    validate, span = reference_validate()
    keep, recs = {}, []
    for case in FLIP_CASES:
        recs.append(run_case(validate, ref_eval, case, keep))
        print(recs[-1]["name"], "attempt", recs[-1]["attempt"], "gap %.2e" % recs[-1]["min_relative_gap"], "cm_sum", recs[-1]["cm_sum"])
    meta = {"source": f"tools/evaluate.py:{span[0]}-{span[1]} (Evaluater.validate, executed unmodified, --flip)",
            "torch": torch.__version__, "numpy": np.__version__, "cases": recs}
    with open(os.path.join(OUT, "flip_kats.json"), "w") as f:
        json.dump(meta, f, indent=1)
    np.savez_compressed(os.path.join(OUT, "flip_tensors.npz"), **keep)
    print("wrote", len(recs), "cases")


if __name__ == "__main__":
    main()
