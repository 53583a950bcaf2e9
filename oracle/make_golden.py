"""Generate ``tests/golden/*`` by running the UNMODIFIED reference
(``/root/reference/utils/loss.py`` and ``utils/eval.py``) in the build container.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden``
The reference cannot travel to the GPU box, so the vectors it produces are
committed; this script is the recipe.  The reference has no tests or golden
vectors of its own, so these ARE the pin.

The reference is loaded by file path because its package ``__init__`` files
import every sibling module, one of which needs ``imageio`` (absent here);
``utils/eval.py:6`` needs ``datasets.cityscapes_Dataset.name_classes``.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

REF = os.environ.get("MSQ_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from maxsquareloss_b200 import synth  # noqa: E402


def load_reference():
    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    ref_loss = by_path("ref_loss", "utils/loss.py")
    sys.modules.setdefault("imageio", types.ModuleType("imageio"))
    # only ``name_classes`` is needed from the datasets package; avoid its __init__
    pkg = types.ModuleType("datasets")
    pkg.__path__ = [os.path.join(REF, "datasets")]
    sys.modules["datasets"] = pkg
    sys.modules["datasets.cityscapes_Dataset"] = by_path("datasets.cityscapes_Dataset",
                                                         "datasets/cityscapes_Dataset.py")
    ref_eval = by_path("ref_eval", "utils/eval.py")
    return ref_loss, ref_eval


def sha(t):
    a = t.numpy() if isinstance(t, torch.Tensor) else np.ascontiguousarray(t)
    return hashlib.sha256(a.tobytes()).hexdigest()


LOSS_CASES = [
    # name, kind, N, shape key, seed, scale, class_bias, quantize, ratio, grad_scale
    ("KAT1_iw_c19_flat", "iw", 1, "cityscapes_target", 0, 1.0, False, False, 0.2, 1.0),
    ("KAT2_iw_c19_peaky", "iw", 1, "cityscapes_target", 1, 5.0, False, False, 0.2, 1.0),
    ("KAT3_ms_c19_n2", "ms", 2, "cityscapes_target", 0, 1.0, False, False, 0.2, 1.0),
    ("KAT4_iw_c16", "iw", 1, "synthia_source", 0, 1.0, False, False, 0.2, 1.0),
    ("KAT5_iw_c13_tiny", "iw", 1, "tiny13", 3, 1.0, False, False, 0.2, 1.0),
    ("iw_c19_biased_scaled", "iw", 1, "cityscapes_target", 2, 5.0, True, False, 0.2, 0.09),
    ("iw_c19_n2_permean", "iw", 2, "cityscapes_target", 5, 3.0, True, False, 0.2, 0.1),
    ("iw_c19_ratio05", "iw", 1, "cityscapes_target", 4, 2.0, False, False, 0.5, 1.0),
    ("iw_c19_dyadic_quant", "iw", 1, "dyadic", 0, 5.0, False, True, 0.2, 1.0),
    ("ms_c19_dyadic_quant", "ms", 1, "dyadic", 1, 5.0, False, True, 0.2, 1.0),
    ("ms_c13_tiny", "ms", 2, "tiny13", 7, 2.0, False, False, 0.2, 0.1),
    ("iw_c19_gta5_shape", "iw", 1, "gta5_source", 3, 4.0, True, False, 0.2, 1.0),
]


def run_loss_case(ref_loss, case, keep_tensors=False):
    name, kind, N, key, seed, scale, bias, quant, ratio, gscale = case
    C, hw, HW = synth.SHAPES[key]
    lo = synth.head_logits(N, C, hw, seed, scale, bias, quant)
    x = lo.clone().requires_grad_(True)
    pred = F.interpolate(x, size=HW, mode='bilinear', align_corners=True)
    prob = F.softmax(pred, dim=1)
    prob.retain_grad()
    if kind == "ms":
        loss = ref_loss.MaxSquareloss(-1, C)(pred, prob)
        hist = None
    else:
        crit = ref_loss.IW_MaxSquareloss(-1, C, ratio)
        # the reference only supports N == 1 (utils/loss.py:98-100); N > 1 is
        # DEFINED as the mean over images of the N == 1 loss
        loss = sum(crit(pred[i:i + 1], prob[i:i + 1]) for i in range(N)) / N
        arg = torch.max(prob.detach(), 1)[1]
        hist = [np.bincount(arg[i].reshape(-1).numpy(), minlength=C).tolist() for i in range(N)]
    (gscale * loss).backward()
    g = x.grad
    z = pred.detach()
    top2 = z.topk(2, 1).values
    gap = top2[:, 0] - top2[:, 1]
    rec = dict(name=name, kind=kind, N=N, shape=key, C=C, hw=list(hw), HW=list(HW), seed=seed,
               scale=scale, class_bias=bias, quantize=quant, ratio=ratio, grad_scale=gscale,
               input_sha256=sha(lo), loss=float(loss.item()), loss_hex=float(loss.item()).hex(),
               grad_sum_abs=float(g.abs().sum().item()), grad_l2=float(g.norm().item()),
               grad_max_abs=float(g.abs().max().item()), hist=hist,
               min_top2_gap=float(gap.min().item()), exact_ties=int((gap == 0).sum().item()),
               argmax_prob_vs_logits_mismatch=int((z.argmax(1) != torch.max(prob.detach(), 1)[1]).sum().item()),
               interp_sha256=sha(z))
    if keep_tensors:
        rec["_tensors"] = dict(logits=lo.numpy(), grad_logits=g.numpy())
        if N == 1:      # full-resolution dL/dprob only for the single-image tiny case (size)
            rec["_tensors"]["grad_prob"] = prob.grad.numpy()
    return rec


def eval_cases(ref_eval):
    out = {}
    meta = []

    def run(name, C, gt, pr, calls=1):
        ev = ref_eval.Eval(C)
        for _ in range(calls):
            ev.add_batch(gt, pr)
        m = dict(name=name, C=C, cm_sum=float(ev.confusion_matrix.sum()),
                 MIoU=ev.Mean_Intersection_over_Union(), PA=ev.Pixel_Accuracy(),
                 MPA=ev.Mean_Pixel_Accuracy(), PC=ev.Mean_Precision(),
                 FWIoU=ev.Frequency_Weighted_Intersection_over_Union())
        if C == 19:
            m["MIoU_16_13"] = ev.Mean_Intersection_over_Union(out_16_13=True)
            m["MPA_16_13"] = ev.Mean_Pixel_Accuracy(out_16_13=True)
            m["PC_16_13"] = ev.Mean_Precision(out_16_13=True)
            m["FWIoU_16_13"] = ev.Frequency_Weighted_Intersection_over_Union(out_16_13=True)
        for k, v in list(m.items()):
            if isinstance(v, tuple):
                m[k] = [float(a) for a in v]
            elif isinstance(v, (np.floating, float)) and k not in ("name",):
                m[k] = float(v)
        meta.append(m)
        out[name + "_gt"] = gt
        out[name + "_pr"] = pr
        out[name + "_cm"] = ev.confusion_matrix.astype(np.int64)

    gt, pr = synth.eval_pair_np((2, 64, 128), 19, 7)            # KAT6
    run("KAT6_c19", 19, gt, pr)
    gt, pr = synth.eval_pair_np((1, 64, 128), 16, 8)            # KAT7
    run("KAT7_c16", 16, gt, pr)
    gt, pr = synth.eval_pair_np((1, 32, 64), 13, 9)
    run("c13_twice", 13, gt, pr, calls=2)
    # ignore value 255 and -1, float ground truth (datasets emit float32 labels)
    rng = np.random.default_rng(11)
    gt = rng.integers(0, 19, (1, 48, 96))
    gt[rng.random(gt.shape) < 0.2] = 255
    gt[rng.random(gt.shape) < 0.1] = -1
    pr = rng.integers(0, 19, gt.shape)
    run("c19_ignore255", 19, gt, pr)
    run("c19_float_gt", 19, gt.astype(np.float32), pr)
    # blocky segmentation-like pair with 30 % noise, tiny (cfg 4 generator)
    g = synth.blocky_labels(1, (64, 128), 16, 1000, grid=(4, 8)).numpy()
    p = synth.noisy_prediction(torch.from_numpy(g), 16, 1000).numpy()
    run("c16_blocky", 16, g, p)
    # all ground truth ignored -> zero matrix -> NaN metrics
    run("c19_all_ignored", 19, np.full((1, 8, 16), -1), np.zeros((1, 8, 16), dtype=np.int64))
    # prediction == C aliases into the next row (reference quirk, utils/eval.py:112)
    gt = np.array([[[0, 1, 2, 3]]]); pr = np.array([[[19, 0, 19, 5]]])
    run("c19_pred_eq_C_alias", 19, gt, pr)
    return out, meta


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    ref_loss, ref_eval = load_reference()
    recs = []
    tensors = {}
    for case in LOSS_CASES:
        small = synth.SHAPES[case[3]][2][0] <= 64
        r = run_loss_case(ref_loss, case, keep_tensors=small)
        t = r.pop("_tensors", None)
        if t is not None:
            for k, v in t.items():
                tensors[f"{r['name']}__{k}"] = v
        recs.append(r)
        print(r["name"], r["loss"], r["grad_sum_abs"], "ties", r["exact_ties"],
              "p-vs-z argmax mismatch", r["argmax_prob_vs_logits_mismatch"])

    # label= argument: histogram from an external label map, gather by argmax (utils/loss.py:87-96)
    C, hw, HW = synth.SHAPES["tiny13"]
    lo = synth.head_logits(1, C, hw, 21, 2.0)
    x = lo.clone().requires_grad_(True)
    pred = F.interpolate(x, size=HW, mode='bilinear', align_corners=True)
    prob = F.softmax(pred, dim=1)
    prob.retain_grad()
    lab = synth.random_labels(1, HW, C, 22)
    loss = ref_loss.IW_MaxSquareloss(-1, C, 0.2)(pred, prob, lab)
    loss.backward()
    tensors["label_arg__logits"] = lo.numpy()
    tensors["label_arg__label"] = lab.numpy()
    tensors["label_arg__grad_logits"] = x.grad.numpy()
    tensors["label_arg__grad_prob"] = prob.grad.numpy()
    recs.append(dict(name="label_arg", kind="iw_label", N=1, shape="tiny13", C=C, hw=list(hw), HW=list(HW),
                     seed=21, scale=2.0, ratio=0.2, label_seed=22, loss=float(loss.item()),
                     grad_sum_abs=float(x.grad.abs().sum().item()), input_sha256=sha(lo),
                     hist=[np.bincount(lab[0].reshape(-1).numpy()[lab[0].reshape(-1).numpy() >= 0],
                                       minlength=C).tolist()]))

    # reference error behaviour that the port/boundary re-define or keep
    errs = {}
    try:
        p2 = torch.softmax(torch.randn(2, 19, 8, 8), 1)
        ref_loss.IW_MaxSquareloss(-1, 19, 0.2)(p2, p2)
        errs["iw_n2"] = "no error"
    except RuntimeError as e:
        errs["iw_n2"] = "RuntimeError: " + str(e)[:80]
    ev = ref_eval.Eval(19)
    try:
        ev.add_batch(np.zeros((1, 4, 4), dtype=np.int64), np.zeros((1, 4, 5), dtype=np.int64))
        errs["eval_shape"] = "no error"
    except AssertionError:
        errs["eval_shape"] = "AssertionError"
    try:
        ev.add_batch(np.zeros((1, 2, 2), dtype=np.int64), -np.ones((1, 2, 2), dtype=np.int64))
        errs["eval_negative_pred"] = "no error"
    except ValueError as e:
        errs["eval_negative_pred"] = "ValueError: " + str(e)[:80]
    try:
        ev.add_batch(np.full((1, 2, 2), 18, dtype=np.int64), np.full((1, 2, 2), 19, dtype=np.int64))
        errs["eval_index_past_end"] = "no error"
    except ValueError as e:
        errs["eval_index_past_end"] = "ValueError: " + str(e)[:80]

    ev_arrays, ev_meta = eval_cases(ref_eval)
    with open(os.path.join(OUT, "loss_kats.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, numpy=np.__version__, cases=recs, reference_errors=errs),
                  f, indent=1)
    with open(os.path.join(OUT, "eval_kats.json"), "w") as f:
        json.dump(dict(numpy=np.__version__, cases=ev_meta), f, indent=1)
    np.savez_compressed(os.path.join(OUT, "loss_tensors.npz"), **tensors)
    np.savez_compressed(os.path.join(OUT, "eval_tensors.npz"), **ev_arrays)
    print("reference error behaviour:", errs)
    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
