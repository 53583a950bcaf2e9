"""Generate ``tests/golden/multi_*`` by executing the reference's OWN ``train_target`` source
(``UDATrainer.train_target``, ``/root/reference/tools/solve_gta5.py:178-218``) with ``--multi``.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_multi``

``tools/solve_gta5.py`` cannot be imported here (``distutils`` is gone from Python 3.12,
``tensorboardX`` / ``imageio`` are absent, module import builds datasets), so the method's source
text is cut out of the file with ``ast`` and compiled unmodified; it runs against a stub ``self``
that carries exactly the attributes the method reads (``args``, ``threshold``, ``device``,
``ignore_index``, ``target_loss`` = the reference's own loss class, ``target_hard_loss`` =
``nn.CrossEntropyLoss(ignore_index=-1)`` as at ``solve_gta5.py:167``).
"""
import ast
import json
import os
import sys
import textwrap
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .make_golden import OUT, REF, load_reference, sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import synth  # noqa: E402


def reference_train_target():
    path = os.path.join(REF, "tools", "solve_gta5.py")
    src = open(path).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "UDATrainer")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "train_target")
    lines = src.splitlines()[fn.lineno - 1:fn.end_lineno]
    ns = {"torch": torch, "F": F, "nn": nn}
    exec(compile(textwrap.dedent("\n".join(lines)), path, "exec"), ns)
    return ns["train_target"], (fn.lineno, fn.end_lineno)


MULTI_CASES = [
    # name, kind, N, shape key, seed, scale, class_bias, threshold, lambda_target, lambda_seg, ratio
    ("multi_iw_c13_tiny", "iw", 1, "tiny13", 31, 4.0, False, 0.95, 0.1, 0.1, 0.2),
    ("multi_ms_c13_tiny_n2", "ms", 2, "tiny13", 32, 4.0, True, 0.9, 0.09, 0.1, 0.2),
    ("multi_iw_c19_cityscapes", "iw", 1, "cityscapes_target", 33, 5.0, False, 0.95, 0.09, 0.1, 0.2),
    ("multi_iw_c19_readme_shape", "iw", 1, "multi_readme", 34, 5.0, True, 0.95, 0.1, 0.1, 0.2),
    ("multi_ms_c19_n2_thr98", "ms", 2, "cityscapes_target", 35, 6.0, True, 0.98, 0.1, 0.1, 0.2),
    ("multi_iw_c16_synthia", "iw", 1, "synthia_source", 36, 4.0, False, 0.95, 0.1, 0.1, 0.2),
    ("multi_iw_c13_none_valid", "iw", 1, "tiny13", 37, 0.5, False, 0.95, 0.1, 0.1, 0.2),
]


def run_case(train_target, ref_loss, case, keep):
    name, kind, N, key, seed, scale, bias, thr, lam_t, lam_s, ratio = case
    C, hw, HW = synth.SHAPES[key]
    lo1 = synth.head_logits(N, C, hw, seed, scale, bias)
    lo2 = synth.second_head(lo1, seed)
    x1, x2 = lo1.clone().requires_grad_(True), lo2.clone().requires_grad_(True)
    # the model's two upsamples (graphs/models/deeplab_multi.py:124,128)
    pred = F.interpolate(x1, size=HW, mode='bilinear', align_corners=True)
    pred_2 = F.interpolate(x2, size=HW, mode='bilinear', align_corners=True)
    stub = types.SimpleNamespace(
        args=types.SimpleNamespace(target_mode="IW_maxsquare" if kind == "iw" else "maxsquare", multi=True,
                                   lambda_target=lam_t, lambda_seg=lam_s),
        threshold=thr, device=torch.device("cpu"), ignore_index=-1,
        target_loss=(ref_loss.IW_MaxSquareloss(-1, C, ratio) if kind == "iw" else ref_loss.MaxSquareloss(-1, C)),
        target_hard_loss=nn.CrossEntropyLoss(ignore_index=-1),
        loss_target_value=0.0, loss_target_value_2=0.0, iter_num=1)
    train_target(stub, (pred, pred_2))
    # label_2 is a local of the method: recompute it the same way for the record, and check it
    # reproduces the method's own loss_target_2
    with torch.no_grad():
        p1, p2 = F.softmax(pred, 1), F.softmax(pred_2, 1)
        lab = torch.where(stub.mask, torch.max((p1 + p2) / 2, 1)[1], torch.ones(1, dtype=torch.long) * -1)
        chk = lam_s * lam_t * F.cross_entropy(pred_2, lab, ignore_index=-1)
        same = torch.equal(chk, stub.loss_target_2.detach()) or (torch.isnan(chk) and torch.isnan(stub.loss_target_2))
        assert same, (chk, stub.loss_target_2)
    # NB the method does `loss_target_ = self.loss_target; loss_target_ += self.loss_target_2`
    # (solve_gta5.py:201,214): an in-place add on the same tensor, so after the call
    # self.loss_target holds the TOTAL.  The head-1 term alone is re-evaluated here.
    with torch.no_grad():
        own = stub.args.lambda_target * stub.target_loss(pred, F.softmax(pred, 1)) if N == 1 or kind == "ms" else None
    g1, g2 = x1.grad, x2.grad
    rec = dict(name=name, kind=kind, N=N, shape=key, C=C, hw=list(hw), HW=list(HW), seed=seed, scale=scale,
               class_bias=bias, threshold=thr, lambda_target=lam_t, lambda_seg=lam_s, ratio=ratio,
               input1_sha256=sha(lo1), input2_sha256=sha(lo2),
               loss_total=float(stub.loss_target.item()), loss_target=float(own.item()),
               loss_target_2=float(stub.loss_target_2.item()),
               nvalid=int(stub.mask.sum().item()), label2_sha256=sha(lab),
               label2_hist=np.bincount(lab.reshape(-1).numpy() + 1, minlength=C + 1).tolist(),
               grad1_sum_abs=float(g1.abs().sum().item()), grad1_l2=float(g1.norm().item()),
               grad2_sum_abs=float(g2.abs().sum().item()), grad2_l2=float(g2.norm().item()))
    t = None
    if keep:
        t = dict(logits1=lo1.numpy(), logits2=lo2.numpy(), label2=lab.numpy().astype(np.int8),
                 grad1=g1.numpy(), grad2=g2.numpy())
    return rec, t


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref_loss, _ = load_reference()
    train_target, span = reference_train_target()
    recs, tensors = [], {}
    for case in MULTI_CASES:
        small = synth.SHAPES[case[3]][2][0] <= 64
        r, t = run_case(train_target, ref_loss, case, small)
        recs.append(r)
        if t:
            for k, v in t.items():
                tensors[f"{r['name']}__{k}"] = v
        print(r["name"], r["loss_total"], r["loss_target"], r["loss_target_2"], "valid", r["nvalid"])
    with open(os.path.join(OUT, "multi_kats.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, source="tools/solve_gta5.py:%d-%d" % span, cases=recs), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "multi_tensors.npz"), **tensors)


if __name__ == "__main__":
    main()
