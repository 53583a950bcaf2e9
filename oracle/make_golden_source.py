"""Generate ``tests/golden/source_*``: the source-side step of the reference's training loop
(``/root/reference/tools/train_source.py:254,280-283``) -- ``nn.CrossEntropyLoss(ignore_index=-1)`` on the
model's upsampled logits, ``np.argmax`` and the reference's OWN ``Eval.add_batch`` (loaded from
``/root/reference/utils/eval.py``) -- run in the build container on seeded synthetic inputs.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_source``
"""
import json
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .make_golden import OUT, load_reference, sha

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import synth  # noqa: E402

SOURCE_CASES = [
    # name, N, shape key, seed, scale, label kind
    ("src_c13_tiny", 2, "tiny13", 41, 2.0, "random"),
    ("src_c13_tiny_blocky", 1, "tiny13", 42, 3.0, "blocky"),
    ("src_c19_gta5", 1, "gta5_source", 43, 3.0, "blocky"),
    ("src_c19_cityscapes_n2", 2, "cityscapes_target", 44, 1.0, "blocky"),
    ("src_c16_synthia", 1, "synthia_source", 45, 2.0, "random"),
    ("src_c13_all_ignored", 1, "tiny13", 46, 1.0, "ignored"),
    ("src_c19_aligned", 1, "cityscapes_target", 47, 1.0, "aligned"),
    ("src_c13_tiny_aligned", 2, "tiny13", 48, 1.0, "aligned"),
]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    _, ref_eval = load_reference()
    crit = nn.CrossEntropyLoss(weight=None, ignore_index=-1)      # tools/train_source.py:128
    recs, tensors = [], {}
    for name, N, key, seed, scale, kind in SOURCE_CASES:
        C, hw, HW = synth.SHAPES[key]
        lo, y = synth.source_case(N, C, hw, HW, seed, scale, kind)
        x = lo.clone().requires_grad_(True)
        pred = F.interpolate(x, size=HW, mode='bilinear', align_corners=True)      # deeplab_multi.py:128
        cur_loss = crit(pred, y)                                                     # train_source.py:254
        (0.5 * cur_loss).backward()
        p = pred.data.cpu().numpy()                                                  # train_source.py:280-283
        label = y.cpu().numpy()
        argpred = np.argmax(p, axis=1)
        ev = ref_eval.Eval(C)
        ev.add_batch(label, argpred)
        g = x.grad
        with np.errstate(all='ignore'):
            miou = ev.Mean_Intersection_over_Union()
        rec = dict(name=name, N=N, shape=key, C=C, hw=list(hw), HW=list(HW), seed=seed, scale=scale, label_kind=kind,
                   grad_scale=0.5, input_sha256=sha(lo), label_sha256=sha(y), loss=float(cur_loss.item()),
                   nvalid=int((y != -1).sum()), grad_sum_abs=float(g.abs().sum().item()), grad_l2=float(g.norm().item()),
                   argpred_sha256=sha(argpred.astype(np.int64)), cm_sha256=sha(ev.confusion_matrix.astype(np.int64)),
                   cm_sum=float(ev.confusion_matrix.sum()), cm_trace=float(np.trace(ev.confusion_matrix)),
                   MIoU=[float(v) for v in miou] if isinstance(miou, tuple) else float(miou))
        recs.append(rec)
        print(name, rec["loss"], rec["nvalid"], rec["cm_trace"], rec["MIoU"])
        if HW[0] <= 64:
            tensors[name + "__logits"] = lo.numpy()
            tensors[name + "__label"] = y.numpy().astype(np.int8)
            tensors[name + "__grad"] = g.numpy()
            tensors[name + "__cm"] = ev.confusion_matrix.astype(np.int64)
    with open(os.path.join(OUT, "source_kats.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, numpy=np.__version__, cases=recs), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "source_tensors.npz"), **tensors)


if __name__ == "__main__":
    main()
