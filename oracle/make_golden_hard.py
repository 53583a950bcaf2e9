"""Generate ``tests/golden/hard_kats.json`` by executing the reference's OWN ``train_target`` source
(``UDATrainer.train_target``, ``/root/reference/tools/solve_gta5.py:178-218``) with ``--target_mode hard`` (no --multi):
pseudo-labels ``argmax(softmax(pred))`` where the maximum probability exceeds the threshold, -1 elsewhere, and
``lambda_target * nn.CrossEntropyLoss(ignore_index=-1)(pred, label)`` (``solve_gta5.py:149-150,185-199``).

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_hard``"""
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .make_golden import OUT, sha
from .make_golden_multi import reference_train_target

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import synth  # noqa: E402

HARD_CASES = [
    # name, N, shape key, seed, scale, class_bias, threshold, lambda_target
    ("hard_c13_tiny", 1, "tiny13", 41, 4.0, False, 0.95, 0.1),
    ("hard_c13_tiny_n2", 2, "tiny13", 42, 4.0, True, 0.9, 0.09),
    ("hard_c19_cityscapes", 1, "cityscapes_target", 43, 5.0, False, 0.95, 0.1),
    ("hard_c16_synthia_thr98", 1, "synthia_source", 44, 6.0, True, 0.98, 0.1),
    ("hard_c13_none_valid", 1, "tiny13", 45, 0.5, False, 0.95, 0.1),
]


def run_case(train_target, case, keep):
    name, N, key, seed, scale, bias, thr, lam_t = case
    C, hw, HW = synth.SHAPES[key]
    lo = synth.head_logits(N, C, hw, seed, scale, bias)
    x = lo.clone().requires_grad_(True)
    pred = F.interpolate(x, size=HW, mode='bilinear', align_corners=True)       # graphs/models/deeplab_multi.py:124
    stub = types.SimpleNamespace(
        args=types.SimpleNamespace(target_mode="hard", multi=False, lambda_target=lam_t, lambda_seg=0.1),
        threshold=thr, device=torch.device("cpu"), ignore_index=-1,
        target_loss=nn.CrossEntropyLoss(ignore_index=-1), target_hard_loss=nn.CrossEntropyLoss(ignore_index=-1),
        loss_target_value=0.0, loss_target_value_2=0.0, iter_num=1)
    train_target(stub, pred)
    with torch.no_grad():       # `label` is a local of the method: recomputed the same way for the record, and checked
        p = F.softmax(pred, 1)
        mx, arg = torch.max(p, 1)
        lab = torch.where(mx > thr, arg, torch.ones(1, dtype=torch.long) * -1)
        chk = lam_t * F.cross_entropy(pred, lab, ignore_index=-1)
        assert torch.equal(chk, stub.loss_target.detach()) or (torch.isnan(chk) and torch.isnan(stub.loss_target))
    g = x.grad
    rec = dict(name=name, N=N, shape=key, C=C, hw=list(hw), HW=list(HW), seed=seed, scale=scale, class_bias=bias,
               threshold=thr, lambda_target=lam_t, input_sha256=sha(lo), loss_target=float(stub.loss_target.item()),
               nvalid=int((lab >= 0).sum().item()), label_sha256=sha(lab),
               label_hist=np.bincount(lab.reshape(-1).numpy() + 1, minlength=C + 1).tolist(),
               grad_sum_abs=float(g.abs().sum().item()), grad_l2=float(g.norm().item()))
    t = dict(logits=lo.numpy(), label=lab.numpy().astype(np.int8), grad=g.numpy()) if keep else None
    return rec, t


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    train_target, span = reference_train_target()
    recs, tensors = [], {}
    for case in HARD_CASES:
        small = synth.SHAPES[case[2]][2][0] <= 64
        r, t = run_case(train_target, case, small)
        recs.append(r)
        if t:
            for k, v in t.items():
                tensors[f"{r['name']}__{k}"] = v
        print(r["name"], r["loss_target"], "valid", r["nvalid"])
    with open(os.path.join(OUT, "hard_kats.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, source="tools/solve_gta5.py:%d-%d" % span, cases=recs), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "hard_tensors.npz"), **tensors)


if __name__ == "__main__":
    main()
