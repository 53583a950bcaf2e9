"""Generate ``tests/golden/perimage_*`` by executing the reference's OWN per-image evaluation loop,
``resultEvaluater.getvalResult`` (``/root/reference/tools/analysis.py:171-240``): per validation image
``Eval.add_batch`` + ``totalEval.add_batch``, the four metrics of the image (``val_info``), ``Eval.reset()``.

TEST INFRASTRUCTURE ONLY.  Run from the repo root:  ``python -m oracle.make_golden_perimage``

``tools/analysis.py`` cannot be imported here (tensorboardX, dataset construction at import), so the method's source
is cut out of the file with ``ast`` and compiled; it runs against a stub ``self`` (the reference's own ``Eval`` twice, a
small recording model, a list as the loader).  Two facts about the fork, both kept:
  * the loop unpacks ``MIoU, IOUS = Eval.Mean_Intersection_over_Union()``, which only works where that method returns a
    pair, i.e. ``num_class == 16`` (``utils/eval.py:49-52``); with 19 classes the reference raises TypeError.  The cases
    are therefore 16-class; MPA and FWIoU come back as (16-class, 13-class) pairs, MIoU is the 16-class value;
  * the statements after the loop (``for key in totalIous``, analysis.py:232-236) iterate over a float and raise: the
    function is compiled WITHOUT that tail (the ``with`` block is cut after the ``for`` loop) and returns ``id2mIOU``.
The image-saving branch (``if MIoU < threshold``) is not entered: threshold = -1.
"""
import ast
import contextlib
import io
import json
import logging
import os
import sys
import types

import numpy as np
import torch
from tqdm import tqdm

from .make_golden import OUT, REF, load_reference, sha
from .make_golden_flip import RecordingModel

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from maxsquareloss_b200 import synth  # noqa: E402


def reference_getval():
    path = os.path.join(REF, "tools", "analysis.py")
    src = open(path).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "resultEvaluater")
    fn = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "getvalResult")
    w = next(n for n in fn.body if isinstance(n, ast.With))
    loop_at = next(i for i, n in enumerate(w.body) if isinstance(n, ast.For))
    w.body = w.body[:loop_at + 1]                       # drop the broken tail after the loop (see the module docstring)
    fn.body = fn.body[:fn.body.index(w) + 1] + [ast.Return(value=ast.Name(id="id2mIOU", ctx=ast.Load()))]
    mod = ast.Module(body=[fn], type_ignores=[])
    ast.fix_missing_locations(mod)
    ns = {"torch": torch, "np": np, "tqdm": tqdm, "os": os, "Path": __import__("pathlib").Path}
    exec(compile(mod, path, "exec"), ns)
    return ns["getvalResult"], (fn.lineno, fn.end_lineno)


PERIMAGE_CASES = [
    # name, images, (H, W), seed, noise
    ("perimage_c16_a", 5, (24, 48), 71, 0.3),
    ("perimage_c16_odd", 4, (19, 33), 72, 0.6),
]


def run_case(getval, ref_eval, case, keep):
    name, K, (H, W), seed, noise = case
    C = 16
    g = torch.Generator().manual_seed(seed)
    batches = []
    for b in range(K):
        x = torch.randn(1, 3, H, W, generator=g)
        y = synth.blocky_labels(1, (H, W), C, seed + b, grid=(4, 8)).unsqueeze(1).float()    # datasets emit float (N,1,H,W)
        batches.append((x, y, [f"img{b}.png"]))
    model = RecordingModel(C, seed, 0.5)
    stub = types.SimpleNamespace(
        logger=logging.getLogger("golden_perimage"), Eval=ref_eval.Eval(C), totalEval=ref_eval.Eval(C), cuda=False,
        device=torch.device("cpu"), val_loader=batches, valid_iterations=K, model=model,
        args=types.SimpleNamespace(show_num_images=1, numpy_transform=False))
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
        id2miou = getval(stub, "/nonexistent", -1.0)
    rows = []
    for (img, pa, mpa, miou, fw), pred, (x, y, _) in zip(id2miou, model.outputs, batches):
        rows.append({"id": img, "PA": float(pa), "MPA": [float(v) for v in mpa], "MIoU": float(miou),
                     "FWIoU": [float(v) for v in fw]})
    for b, (pred, (x, y, _)) in enumerate(zip(model.outputs, batches)):
        keep[f"{name}/pred{b}"] = pred.numpy()
        keep[f"{name}/label{b}"] = y.squeeze(1).long().numpy().astype(np.int16)
    tot = stub.totalEval
    keep[f"{name}/total_cm"] = tot.confusion_matrix.astype(np.int64)
    return {"name": name, "C": C, "images": K, "HW": [H, W], "per_image": rows,
            "total_cm_sha": sha(tot.confusion_matrix.astype(np.int64)), "total_cm_sum": int(tot.confusion_matrix.sum()),
            "total_miou": [float(v) for v in tot.Mean_Intersection_over_Union()]}


def main():
    logging.getLogger("golden_perimage").addHandler(logging.NullHandler())
    logging.getLogger("golden_perimage").propagate = False
    _, ref_eval = load_reference()
    getval, span = reference_getval()
    keep, recs = {}, []
    for case in PERIMAGE_CASES:
        recs.append(run_case(getval, ref_eval, case, keep))
        print(recs[-1]["name"], [round(r["MIoU"], 4) for r in recs[-1]["per_image"]], recs[-1]["total_cm_sum"])
    meta = {"source": f"tools/analysis.py:{span[0]}-{span[1]} (resultEvaluater.getvalResult, the per-image loop executed "
                      "unmodified; the statements after the loop cut, see oracle/make_golden_perimage.py)",
            "torch": torch.__version__, "numpy": np.__version__, "cases": recs}
    with open(os.path.join(OUT, "perimage_kats.json"), "w") as f:
        json.dump(meta, f, indent=1)
    np.savez_compressed(os.path.join(OUT, "perimage_tensors.npz"), **keep)


if __name__ == "__main__":
    main()
