"""Generate ``tests/golden/entropy_*`` by running the UNMODIFIED reference classes ``softCrossEntropy`` and
``IWsoftCrossEntropy`` (``/root/reference/utils/loss.py:17-67``) the way the trainers call them
(``tools/solve_gta5.py:183,188-190,199``: ``target_loss(pred, softmax(pred))``, target attached).

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_entropy``
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

ENTROPY_CASES = [
    # name, iw, N, shape key, seed, scale, class_bias, ratio, grad_scale
    ("ent_c13_tiny_n2", False, 2, "tiny13", 51, 2.0, False, 0.2, 0.1),
    ("iwent_c13_tiny", True, 1, "tiny13", 52, 2.0, False, 0.2, 1.0),
    ("ent_c19_cityscapes_n2", False, 2, "cityscapes_target", 53, 3.0, True, 0.2, 0.1),
    ("iwent_c19_cityscapes", True, 1, "cityscapes_target", 54, 5.0, False, 0.2, 0.09),
    ("iwent_c19_biased_ratio05", True, 1, "cityscapes_target", 55, 2.0, True, 0.5, 1.0),
    ("iwent_c16_synthia", True, 1, "synthia_source", 56, 3.0, False, 0.2, 1.0),
    ("iwent_c19_dyadic_quant", True, 1, "dyadic", 57, 5.0, False, 0.2, 1.0),
]


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref_loss, _ = load_reference()
    recs, tensors = [], {}
    for name, iw, N, key, seed, scale, bias, ratio, gs in ENTROPY_CASES:
        C, hw, HW = synth.SHAPES[key]
        lo = synth.head_logits(N, C, hw, seed, scale, bias, quantize=(key == "dyadic"))
        x = lo.clone().requires_grad_(True)
        pred = F.interpolate(x, size=HW, mode='bilinear', align_corners=True)
        prob = F.softmax(pred, dim=1)
        crit = ref_loss.IWsoftCrossEntropy(-1, C, ratio) if iw else ref_loss.softCrossEntropy(-1)
        loss = crit(pred, prob)
        (gs * loss).backward()
        g = x.grad
        arg = torch.max(pred.detach(), 1)[1]
        hist = [np.bincount(arg[i].reshape(-1).numpy(), minlength=C).tolist() for i in range(N)] if iw else None
        top2 = pred.detach().topk(2, 1).values
        rec = dict(name=name, iw=iw, N=N, shape=key, C=C, hw=list(hw), HW=list(HW), seed=seed, scale=scale, class_bias=bias,
                   quantize=(key == "dyadic"), ratio=ratio, grad_scale=gs, input_sha256=sha(lo), loss=float(loss.item()),
                   grad_sum_abs=float(g.abs().sum().item()), grad_l2=float(g.norm().item()), hist=hist,
                   exact_ties=int((top2[:, 0] == top2[:, 1]).sum().item()))
        recs.append(rec)
        print(name, rec["loss"], rec["grad_sum_abs"], "ties", rec["exact_ties"])
        if HW[0] <= 64:
            tensors[name + "__logits"] = lo.numpy()
            tensors[name + "__grad"] = g.numpy()
    with open(os.path.join(OUT, "entropy_kats.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=recs), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "entropy_tensors.npz"), **tensors)


if __name__ == "__main__":
    main()
