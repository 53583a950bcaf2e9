"""Per-image evaluation (Eval.add_batch_per_image -> msq_confusion_i64_multi / msq_confusion_per_image_logits_f32) and the
deferred add_batch queue (Eval(defer=K)) against vectors frozen from the reference's own per-image loop
(tools/analysis.py:171-240) and against the NumPy port.  Everything here is integer or float64-on-host: bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from maxsquareloss_b200 import synth
from oracle import eval_port

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "perimage_kats.json")) as _f:
    CASES = json.load(_f)["cases"]


@pytest.fixture(scope="module")
def msq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as m
    from maxsquareloss_b200 import _lib, build
    build.build()
    _lib.load()
    return m


def _check_rows(ev, case):
    rows = ev.per_image_metrics()
    assert len(rows) == len(case["per_image"])
    for (pa, mpa, miou, fw), want in zip(rows, case["per_image"]):
        assert pa == want["PA"]
        assert list(mpa) == want["MPA"]
        assert miou[0] == want["MIoU"]               # analysis.py:179 unpacks the (16-class, 13-class) pair
        assert list(fw) == want["FWIoU"]


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
@pytest.mark.parametrize("how", ["argmax_one_call_per_image", "argmax_one_batch", "logits_one_batch"])
def test_per_image_vs_reference_golden(msq, case, how):
    t = np.load(os.path.join(GOLDEN, "perimage_tensors.npz"))
    name, C, K = case["name"], case["C"], case["images"]
    preds = [torch.from_numpy(t[f"{name}/pred{b}"]).cuda() for b in range(K)]
    labels = [torch.from_numpy(t[f"{name}/label{b}"].astype(np.int64)).cuda() for b in range(K)]
    ev = msq.Eval(C)
    if how == "argmax_one_call_per_image":
        for p, y in zip(preds, labels):
            ev.add_batch_per_image(y, torch.from_numpy(np.argmax(p.cpu().numpy(), axis=1)).cuda())
    elif how == "argmax_one_batch":
        ev.add_batch_per_image(torch.cat(labels), torch.from_numpy(np.argmax(torch.cat(preds).cpu().numpy(), axis=1)).cuda())
    else:
        ev.add_batch_per_image(torch.cat(labels), torch.cat(preds))
    _check_rows(ev, case)
    assert np.array_equal(ev.confusion_matrix, t[f"{name}/total_cm"])          # the running total (analysis.py totalEval)
    assert list(ev.Mean_Intersection_over_Union()) == case["total_miou"]
    assert ev.per_image_matrices().sum() == case["total_cm_sum"]
    ev.reset()
    assert ev.per_image_matrices().shape == (0, C, C)


@pytest.mark.parametrize("C", [13, 16, 19, 7])
def test_per_image_ragged_and_unaligned(msq, C):
    """Images of different sizes in one launch, pixel counts not divisible by 4, pointers off 16-byte alignment."""
    rng = np.random.default_rng(C)
    sizes = [(5, 7), (16, 32), (1, 1), (33, 3), (64, 64)]
    ev = msq.Eval(C)
    pairs, want = [], []
    for i, hw in enumerate(sizes):
        gt = rng.integers(-1, C, hw)
        gt[0, 0] = 255
        pr = rng.integers(0, C, hw)
        want.append(eval_port.confusion(gt, pr, C))
        gbuf = torch.zeros(gt.size + 1, dtype=torch.int64, device="cuda")
        pbuf = torch.zeros(gt.size + 1, dtype=torch.int64, device="cuda")
        off = i % 2                                                # every other pair starts 8 bytes off
        gbuf[off:off + gt.size] = torch.from_numpy(gt.reshape(-1)).cuda()
        pbuf[off:off + gt.size] = torch.from_numpy(pr.reshape(-1)).cuda()
        pairs.append((gbuf[off:off + gt.size], pbuf[off:off + gt.size]))
    block = torch.zeros(len(sizes), C * C, dtype=torch.int64, device="cuda")
    ev._launch_multi(pairs, block.data_ptr(), C * C, ev._cm_ptr)
    got = block.cpu().numpy().reshape(-1, C, C)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    assert np.array_equal(ev.confusion_matrix, np.sum(want, axis=0))


def test_more_pairs_than_one_launch_holds(msq):
    C, K = 16, 75                                                      # 32 pairs travel per launch
    gts = [synth.blocky_labels(1, (16, 32), C, i, grid=(4, 8)).cuda() for i in range(K)]
    prs = [synth.noisy_prediction(g.cpu(), C, i).cuda() for i, g in enumerate(gts)]
    ev = msq.Eval(C)
    ev.add_batch_per_image(torch.cat(gts), torch.cat(prs))
    mats = ev.per_image_matrices()
    assert mats.shape == (K, C, C)
    for i in range(K):
        assert np.array_equal(mats[i], eval_port.confusion(gts[i].cpu().numpy(), prs[i].cpu().numpy(), C))


@pytest.mark.parametrize("defer", [1, 4, 16, 40])
def test_deferred_add_batch_equals_eager(msq, defer):
    C, K = 19, 37
    gts = [synth.blocky_labels(1, (32, 64), C, 100 + i, grid=(4, 8)).cuda() for i in range(K)]
    prs = [synth.noisy_prediction(g.cpu(), C, 100 + i).cuda() for i, g in enumerate(gts)]
    eager, lazy = msq.Eval(C), msq.Eval(C, defer=defer)
    port = eval_port.EvalPort(C)
    for g, p in zip(gts, prs):
        eager.add_batch(g, p)
        lazy.add_batch(g, p)
        port.add_batch(g.cpu().numpy(), p.cpu().numpy())
    assert np.array_equal(lazy.confusion_matrix, port.confusion_matrix)
    assert np.array_equal(eager.confusion_matrix, port.confusion_matrix)
    assert lazy.Mean_Intersection_over_Union() == port.Mean_Intersection_over_Union()
    # numpy callers stay synchronous (the reference's behaviour) and flush what is queued
    lazy.add_batch(gts[0], prs[0])
    lazy.add_batch(gts[1].cpu().numpy(), prs[1].cpu().numpy())
    port.add_batch(gts[0].cpu().numpy(), prs[0].cpu().numpy())
    port.add_batch(gts[1].cpu().numpy(), prs[1].cpu().numpy())
    assert np.array_equal(lazy.confusion_matrix, port.confusion_matrix)


def test_deferred_add_batch_detects_in_place_modification(msq):
    C = 13
    g = synth.blocky_labels(1, (16, 32), C, 1, grid=(4, 8)).cuda()
    p = synth.noisy_prediction(g.cpu(), C, 1).cuda()
    ev = msq.Eval(C, defer=8)
    ev.add_batch(g, p)
    p.add_(1)                                                          # the queued launch has not read it yet
    with pytest.raises(RuntimeError):
        ev.confusion_matrix


def test_error_flag_keeps_the_valid_counts(msq):
    """one out-of-contract batch must not discard the batches accumulated before it."""
    C = 13
    g = synth.blocky_labels(1, (16, 32), C, 2, grid=(4, 8)).cuda()
    p = synth.noisy_prediction(g.cpu(), C, 2).cuda()
    ev = msq.Eval(C)
    ev.add_batch(g, p)
    bad = p.clone()
    bad[0, 0, 0] = -500                                                # numpy.bincount would raise ValueError
    g2 = g.clone()
    g2[0, 0, 0] = 0
    ev.add_batch(g2, bad)
    with pytest.raises(ValueError):
        ev.confusion_matrix
    want = eval_port.confusion(g.cpu().numpy(), p.cpu().numpy(), C)
    g2n, badn = g2.cpu().numpy().copy(), bad.cpu().numpy()
    keep = np.ones_like(g2n, dtype=bool)
    keep[0, 0, 0] = False
    want = want + eval_port.confusion(g2n[keep], badn[keep], C)
    assert np.array_equal(ev.confusion_matrix, want)                   # the in-contract pixels are all there
