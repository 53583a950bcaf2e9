"""Generate ``tests/golden/softce_*`` by running the UNMODIFIED reference classes ``softCrossEntropy`` and
``IWsoftCrossEntropy`` (``/root/reference/utils/loss.py:17-67``) on full-resolution ``inputs`` with targets that are
NOT ``softmax(inputs)``: a detached softmax of other logits, a distribution with entries equal to ``ignore_index``
(-1, masked by ``target != ignore_index``), and raw real values.  Both arguments require grad; what is frozen is the
loss and d/d inputs, d/d target.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_softce``
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from .make_golden import OUT, load_reference, sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import synth  # noqa: E402

SOFTCE_CASES = [
    # name, iw, N, C, (H, W), seed, scale, target kind, ratio, grad_scale
    ("sce_c19_other_softmax_n2", False, 2, 19, (24, 48), 61, 3.0, "other_softmax", 0.2, 0.1),
    ("sce_c13_masked", False, 1, 13, (17, 31), 62, 2.0, "masked", 0.2, 1.0),
    ("sce_c7_raw_odd", False, 2, 7, (9, 13), 63, 2.0, "raw", 0.2, 1.0),
    ("iwsce_c19_other_softmax", True, 1, 19, (24, 48), 64, 4.0, "other_softmax", 0.2, 0.09),
    ("iwsce_c16_masked_ratio05", True, 1, 16, (20, 34), 65, 3.0, "masked", 0.5, 1.0),
    ("iwsce_c13_raw", True, 1, 13, (17, 31), 66, 2.0, "raw", 0.2, 1.0),
    ("iwsce_c19_self_softmax", True, 1, 19, (24, 48), 67, 5.0, "self", 0.2, 1.0),
]


def make_inputs(case):
    name, iw, N, C, (H, W), seed, scale, kind, ratio, gs = case
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(N, C, H, W, generator=g) * scale
    if kind == "other_softmax":
        t = F.softmax(torch.randn(N, C, H, W, generator=g) * 2.0, dim=1)
    elif kind == "masked":
        t = F.softmax(torch.randn(N, C, H, W, generator=g) * 2.0, dim=1)
        drop = torch.rand(N, C, H, W, generator=g) < 0.15
        t = torch.where(drop, torch.full_like(t, -1.0), t)           # entries equal to ignore_index are masked out
    elif kind == "raw":
        t = torch.randn(N, C, H, W, generator=g)
    elif kind == "self":
        t = None
    else:
        raise ValueError(kind)
    return z, t


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref_loss, _ = load_reference()
    recs, tensors = [], {}
    for case in SOFTCE_CASES:
        name, iw, N, C, (H, W), seed, scale, kind, ratio, gs = case
        z, t = make_inputs(case)
        x = z.clone().requires_grad_(True)
        if t is None:                                                # the trainers' call: target attached to the graph
            tt = F.softmax(x, dim=1)
            tt.retain_grad()
        else:
            tt = t.clone().requires_grad_(True)
        crit = ref_loss.IWsoftCrossEntropy(-1, C, ratio) if iw else ref_loss.softCrossEntropy(-1)
        loss = crit(x, tt)
        (gs * loss).backward()
        arg = torch.max(z, 1)[1]
        hist = [np.bincount(arg[i].reshape(-1).numpy(), minlength=C).tolist() for i in range(N)] if iw else None
        rec = dict(name=name, iw=iw, N=N, C=C, HW=[H, W], seed=seed, scale=scale, target=kind, ratio=ratio, grad_scale=gs,
                   input_sha256=sha(z), loss=float(loss.item()), hist=hist,
                   grad_inputs_l2=float(x.grad.norm().item()), grad_target_l2=float(tt.grad.norm().item()))
        recs.append(rec)
        tensors[name + "__inputs"] = z.numpy()
        if t is not None:
            tensors[name + "__target"] = t.numpy()
        tensors[name + "__grad_inputs"] = x.grad.numpy()         # "self": the TOTAL derivative (both paths)
        tensors[name + "__grad_target"] = tt.grad.numpy()
        print(name, rec["loss"], rec["grad_inputs_l2"], rec["grad_target_l2"])
    with open(os.path.join(OUT, "softce_kats.json"), "w") as f:
        json.dump(dict(source="utils/loss.py:17-67 (softCrossEntropy, IWsoftCrossEntropy), executed unmodified",
                       torch=torch.__version__, cases=recs), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "softce_tensors.npz"), **tensors)


if __name__ == "__main__":
    main()
