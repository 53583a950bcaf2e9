"""CUDA confusion-matrix kernels (through the C ABI via maxsquareloss_b200.Eval) against the
oracle and the vectors frozen from the reference.  Bit-exact everywhere."""
import warnings

import numpy as np
import pytest
import torch

from maxsquareloss_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def msq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as m
    from maxsquareloss_b200 import _lib, build
    build.build()        # no-op when the in-tree library is current
    _lib.load()          # fail loudly if the extension is missing
    return m


def _same(a, b):
    a = np.atleast_1d(np.asarray(a, dtype=np.float64))
    b = np.atleast_1d(np.asarray(b, dtype=np.float64))
    return np.array_equal(a, b, equal_nan=True)


def test_golden_vectors_numpy_inputs(msq, eval_kats, eval_tensors):
    """The reference callers' calling convention: numpy maps in, float64 matrix + metrics out."""
    for m in eval_kats["cases"]:
        gt, pr, cm = (eval_tensors[m["name"] + s] for s in ("_gt", "_pr", "_cm"))
        ev = msq.Eval(m["C"])
        for _ in range(2 if m["name"].endswith("twice") else 1):
            ev.add_batch(gt, pr)
        assert ev.confusion_matrix.dtype == np.float64
        assert np.array_equal(ev.confusion_matrix.astype(np.int64), cm), m["name"]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert _same(ev.Mean_Intersection_over_Union(), m["MIoU"]), m["name"]        # bit-exact mIoU
            assert _same(ev.Pixel_Accuracy(), m["PA"])
            assert _same(ev.Mean_Pixel_Accuracy(), m["MPA"])
            assert _same(ev.Mean_Precision(), m["PC"])
            assert _same(ev.Frequency_Weighted_Intersection_over_Union(), m["FWIoU"])
            if m["C"] == 19:
                assert _same(ev.Mean_Intersection_over_Union(out_16_13=True), m["MIoU_16_13"])
                assert _same(ev.Mean_Pixel_Accuracy(out_16_13=True), m["MPA_16_13"])
                assert _same(ev.Mean_Precision(out_16_13=True), m["PC_16_13"])
                assert _same(ev.Frequency_Weighted_Intersection_over_Union(out_16_13=True), m["FWIoU_16_13"])
        ev.reset()
        assert ev.confusion_matrix.sum() == 0


@pytest.mark.parametrize("C,shape", [(19, (2, 720, 1280)), (16, (1, 512, 1024)), (13, (3, 37, 53)),
                                     (19, (1, 1, 5)), (32, (1, 64, 64)), (1, (1, 8, 8)), (2, (1, 33, 7))])
@pytest.mark.parametrize("agg", [0, 1, 2])
def test_cuda_inputs_vs_oracle(msq, C, shape, agg):
    from maxsquareloss_b200 import _lib
    from oracle import eval_port
    _lib.tune("conf_agg", agg)
    try:
        for kind in ("blocky", "uniform"):
            if kind == "blocky":
                gt = synth.blocky_labels(shape[0], shape[1:], C, 5, grid=(min(16, shape[1]), min(32, shape[2])))
                pr = synth.noisy_prediction(gt, C, 5)
            else:
                gt = synth.random_labels(shape[0], shape[1:], C, 6)
                gt[0, 0, 0] = 255                       # also ignored
                pr = synth.random_labels(shape[0], shape[1:], C, 7).clamp_(min=0)
            ref = eval_port.confusion(gt.numpy(), pr.numpy(), C)
            ev = msq.Eval(C)
            ev.add_batch(gt.cuda(), pr.cuda())
            assert np.array_equal(ev.confusion_matrix.astype(np.int64), ref), (kind, agg)
    finally:
        _lib.tune("conf_agg", 0)


def test_unaligned_views_and_accumulation(msq):
    from oracle import eval_port
    C = 19
    gt = synth.blocky_labels(1, (97, 131), C, 9)
    pr = synth.noisy_prediction(gt, C, 9)
    g = gt.cuda().reshape(-1)[1:]          # 8-byte but not 16-byte aligned
    p = pr.cuda().reshape(-1)[1:]
    ev = msq.Eval(C)
    ev.add_batch(g, p)
    ev.add_batch(g, p)
    ref = 2 * eval_port.confusion(gt.numpy().reshape(-1)[1:], pr.numpy().reshape(-1)[1:], C)
    assert np.array_equal(ev.confusion_matrix.astype(np.int64), ref)
    ev.add_batch(torch.empty(0, dtype=torch.int64).cuda(), torch.empty(0, dtype=torch.int64).cuda())   # empty batch
    assert np.array_equal(ev.confusion_matrix.astype(np.int64), ref)


@pytest.mark.parametrize("C,shape", [(19, (2, 91 * 4, 161 * 4)), (16, (1, 512, 1024)), (13, (2, 37, 53)),
                                     (5, (1, 64, 64)), (32, (1, 32, 40)), (21, (1, 48, 50))])
def test_fused_argmax_from_logits(msq, C, shape):
    """np.argmax(pred, axis=1) + add_batch (tools/train_source.py:457-459,492) in one kernel."""
    from oracle import eval_port
    g = torch.Generator().manual_seed(3)
    lg = torch.randn(shape[0], C, *shape[1:], generator=g)
    lg[:, :, 0, :8] = torch.round(lg[:, :, 0, :8])          # exact ties: the first maximum must win
    lg[0, 2, 1, 3] = float("nan")                           # NaN is the maximum for numpy.argmax
    gt = synth.blocky_labels(shape[0], shape[1:], C, 4, grid=(min(16, shape[1]), min(32, shape[2])))
    ref = eval_port.confusion(gt.numpy(), np.argmax(lg.numpy(), axis=1), C)
    ev = msq.Eval(C)
    ev.add_batch(gt.cuda(), lg.cuda())
    assert np.array_equal(ev.confusion_matrix.astype(np.int64), ref)
    assert np.array_equal(msq.fast_hist(gt.cuda(), lg.cuda(), C), ref)


def test_error_conventions(msq):
    ev = msq.Eval(19)
    with pytest.raises(AssertionError):                                            # utils/eval.py:119
        ev.add_batch(np.zeros((1, 4, 4), dtype=np.int64), np.zeros((1, 4, 5), dtype=np.int64))
    with pytest.raises(ValueError):                                                # numpy.bincount
        ev.add_batch(np.zeros((1, 2, 2), dtype=np.int64), -np.ones((1, 2, 2), dtype=np.int64))
    ev.reset()
    with pytest.raises(ValueError):                                                # reshape past C*C
        ev.add_batch(np.full((1, 2, 2), 18, dtype=np.int64), np.full((1, 2, 2), 19, dtype=np.int64))
    ev.reset()
    # CUDA inputs: the error is raised when the matrix is next read (no host sync in add_batch)
    ev.add_batch(torch.zeros(4, dtype=torch.int64).cuda(), -torch.ones(4, dtype=torch.int64).cuda())
    with pytest.raises(ValueError):
        _ = ev.confusion_matrix
    with pytest.raises(RuntimeError):
        ev.add_batch_logits(torch.zeros(1, 2, 2, dtype=torch.int64), torch.zeros(1, 19, 2, 2))   # CPU logits


def test_full_size_properties(msq):
    """cfg 4 scale (subset of the 500 val images): size-independent invariants."""
    C, n, hw = 16, 8, (512, 1024)
    gt = synth.blocky_labels(n, hw, C, 1000).cuda()
    pr = synth.noisy_prediction(gt.cpu(), C, 1000).cuda()
    ev = msq.Eval(C)
    ev.add_batch(gt, pr)
    cm = ev.confusion_matrix
    valid = gt >= 0
    assert cm.sum() == int(valid.sum())
    assert np.array_equal(cm.sum(axis=1).astype(np.int64), torch.bincount(gt[valid], minlength=C).cpu().numpy())
    assert np.array_equal(cm.sum(axis=0).astype(np.int64), torch.bincount(pr[valid], minlength=C).cpu().numpy())
    ev2 = msq.Eval(C)                      # additivity over images == one big batch
    for i in range(n):
        ev2.add_batch(gt[i:i + 1], pr[i:i + 1])
    assert np.array_equal(ev2.confusion_matrix, cm)
    miou = ev.Mean_Intersection_over_Union()
    assert isinstance(miou, tuple) and len(miou) == 2          # 16-class Eval returns (mIoU16, mIoU13)


# ------------------------------------------------------------------ flip-ensemble evaluation (tools/evaluate.py:120-141)
@pytest.mark.parametrize("C,shape,scale", [(19, (2, 64, 128), 3.0), (16, (1, 40, 72), 1.0), (13, (1, 33, 65), 5.0),
                                           (7, (2, 9, 31), 2.0), (19, (1, 512, 1024), 4.0), (19, (1, 17, 3), 0.01),
                                           (19, (1, 16, 8), 0.01), (16, (2, 12, 6), 0.02)])
def test_flip_ensemble_vs_oracle(msq, C, shape, scale):
    from oracle import eval_port
    n, h, w = shape
    g = torch.Generator().manual_seed(C * 1000 + w)
    a = torch.randn(n, C, h, w, generator=g) * scale
    b = torch.flip(a, dims=[-1]) + 0.5 * scale * torch.randn(n, C, h, w, generator=g)     # a flipped view of a similar scene
    gt = synth.blocky_labels(n, (h, w), C, 17, grid=(4, 8))
    arg = eval_port.flip_ensemble_argmax(a, b)                    # CPU torch arithmetic
    arg_dev = eval_port.flip_ensemble_argmax(a.cuda(), b.cuda())  # the same ops on this GPU (what the reference runs)
    port = eval_port.EvalPort(C)
    port.add_batch(gt.numpy(), arg_dev)
    ev = msq.Eval(C)
    ev.add_batch_flip(gt.cuda(), a.cuda(), b.cuda())
    assert np.array_equal(ev.confusion_matrix, port.confusion_matrix)
    # CPU and CUDA softmax may round differently: any disagreement must be a pixel whose two best averaged
    # probabilities are within a few ulps
    assert (arg != arg_dev).mean() < 1e-4
    ev2 = msq.Eval(C)
    ev2.add_batch(gt.cuda(), torch.from_numpy(arg_dev).cuda())
    assert np.array_equal(ev.confusion_matrix, ev2.confusion_matrix)
    assert ev.Mean_Intersection_over_Union() == ev2.Mean_Intersection_over_Union() or C == 16


def test_flip_ensemble_identical_views_reduce_to_plain_argmax(msq):
    a = torch.randn(1, 19, 32, 48) * 3
    gt = synth.random_labels(1, (32, 48), 19, 2)
    ev = msq.Eval(19)
    ev.add_batch_flip(gt.cuda(), a.cuda(), torch.flip(a, dims=[-1]).cuda())
    ref = msq.Eval(19)
    ref.add_batch_logits(gt.cuda(), a.cuda())
    assert np.array_equal(ev.confusion_matrix, ref.confusion_matrix)
    with pytest.raises(AssertionError):
        ev.add_batch_flip(gt.cuda(), a.cuda(), a[:, :, :16].cuda())
    with pytest.raises(RuntimeError):
        ev.add_batch_flip(gt, a, a)


def test_flip_ensemble_golden_from_reference_validate(msq):
    """Vectors frozen from the reference's own Evaluater.validate with --flip (oracle/make_golden_flip.py): logits for the
    image and for its mirror image as the model returned them, float labels; confusion matrix and mIoU bit-exact.  No pixel
    of these cases has its two best averaged probabilities within 2e-4 relative."""
    import json
    import os
    GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    with open(os.path.join(GOLDEN, "flip_kats.json")) as f:
        cases = json.load(f)["cases"]
    t = np.load(os.path.join(GOLDEN, "flip_tensors.npz"))
    for case in cases:
        name, C = case["name"], case["C"]
        ev = msq.Eval(C)
        for b in range(case["batches"]):
            ev.add_batch_flip(torch.from_numpy(t[f"{name}/label{b}"].astype(np.int64)).cuda(),
                              torch.from_numpy(t[f"{name}/pred{b}"]).cuda(), torch.from_numpy(t[f"{name}/pred_flip{b}"]).cuda())
        assert np.array_equal(ev.confusion_matrix, t[f"{name}/cm"]), name
        miou = ev.Mean_Intersection_over_Union()
        assert (list(miou) if isinstance(miou, tuple) else miou) == case["miou"], name
