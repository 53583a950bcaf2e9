"""The oracle (oracle/) against the vectors frozen from the reference itself
(tests/golden/, produced by oracle/make_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from maxsquareloss_b200 import synth
from oracle import bilinear, eval_port, loss_math, loss_port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "loss_kats.json")) as _f:
    _CASES = [c for c in json.load(_f)["cases"] if c["kind"] in ("iw", "ms")]
with open(os.path.join(GOLDEN, "eval_kats.json")) as _f:
    _ECASES = json.load(_f)["cases"]


def _logits(c):
    lo = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"], c["quantize"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"], \
        "seeded generator no longer reproduces the frozen input"
    return lo


@pytest.mark.parametrize("shape", [((9, 17), (64, 128)), ((65, 129), (512, 1024)), ((65, 129), (513, 1025)),
                                   ((91, 161), (720, 1280)), ((96, 161), (760, 1280)), ((81, 161), (640, 1280)),
                                   ((33, 65), (33, 65)), ((17, 40), (200, 41))])
def test_bilinear_bit_exact_vs_torch(shape):
    """Every head/label geometry of the BASELINE configs: the restated fp32
    arithmetic equals F.interpolate bit for bit."""
    (h, w), (H, W) = shape
    g = torch.Generator().manual_seed(h * 1000 + W)
    lo = torch.randn(1, 4, h, w, generator=g) * 3
    ref = F.interpolate(lo, size=(H, W), mode='bilinear', align_corners=True).numpy()
    mine = bilinear.upsample(lo.numpy(), (H, W))
    assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))


@pytest.mark.parametrize("shape", [((3, 5), (7, 9)), ((1, 6), (5, 11)), ((5, 3), (1, 9)), ((3, 5), (33, 64))])
def test_bilinear_small_outputs_within_rounding(shape):
    """For small outputs torch's CPU kernel takes a differently-contracted code
    path (last-bit differences vs its own large-tensor path and vs the CUDA
    kernel, whose SASS uses the one formula at every size); the restatement
    follows the CUDA / large-tensor arithmetic, so here only closeness holds."""
    (h, w), (H, W) = shape
    g = torch.Generator().manual_seed(h * 1000 + W)
    lo = torch.randn(2, 3, h, w, generator=g) * 3
    ref = F.interpolate(lo, size=(H, W), mode='bilinear', align_corners=True).numpy()
    mine = bilinear.upsample(lo.numpy(), (H, W))
    assert np.abs(ref - mine).max() <= 1e-6


def test_bilinear_adjoint_is_the_adjoint():
    rng = np.random.default_rng(0)
    lo = rng.standard_normal((1, 2, 5, 7)).astype(np.float32)
    g = rng.standard_normal((1, 2, 19, 23))
    up = bilinear.upsample(lo, (19, 23)).astype(np.float64)
    lhs = (up * g).sum()
    rhs = (lo.astype(np.float64) * bilinear.upsample_adjoint(g, (5, 7))).sum()
    assert abs(lhs - rhs) <= 1e-6 * max(1.0, abs(lhs))


@pytest.mark.parametrize("c", _CASES, ids=[c["name"] for c in _CASES])
def test_port_reproduces_reference(c):
    lo = _logits(c)
    if c["kind"] == "iw":
        loss, grad, hist = loss_port.chain_iw_maxsquare(lo, c["HW"], c["C"], c["ratio"], c["grad_scale"])
        assert hist.tolist() == c["hist"]
    else:
        loss, grad = loss_port.chain_maxsquare(lo, c["HW"], c["grad_scale"])
    # same torch build -> same bits up to the thread-count-dependent reduction order
    assert abs(loss.item() - c["loss"]) <= 1e-6 * abs(c["loss"])
    assert abs(grad.abs().sum().item() - c["grad_sum_abs"]) <= 1e-5 * c["grad_sum_abs"]
    assert abs(grad.norm().item() - c["grad_l2"]) <= 1e-5 * c["grad_l2"]


@pytest.mark.parametrize("c", _CASES, ids=[c["name"] for c in _CASES])
def test_closed_form_matches_reference(c):
    lo = _logits(c).numpy()
    if c["kind"] == "iw":
        r = loss_math.fused_iw(lo, c["HW"], c["C"], c["ratio"], c["grad_scale"])
        assert r["hist"].tolist() == c["hist"]           # bit-exact class histograms
    else:
        r = loss_math.fused_ms(lo, c["HW"], c["grad_scale"])
    assert hashlib.sha256(r["z"].tobytes()).hexdigest() == c["interp_sha256"]   # bit-exact upsample
    assert abs(r["loss"] - c["loss"]) <= 1e-5 * abs(c["loss"])                # north-star tolerance
    g = r["grad_logits"]
    assert abs(np.abs(g).sum() - c["grad_sum_abs"]) <= 1e-4 * c["grad_sum_abs"]
    assert abs(np.sqrt((g * g).sum()) - c["grad_l2"]) <= 1e-4 * c["grad_l2"]


def test_closed_form_gradients_elementwise(loss_tensors):
    for name, kind in (("KAT5_iw_c13_tiny", "iw"), ("ms_c13_tiny", "ms")):
        c = next(x for x in _CASES if x["name"] == name)
        lo = loss_tensors[name + "__logits"]
        ref_g = loss_tensors[name + "__grad_logits"].astype(np.float64)
        r = (loss_math.fused_iw(lo, c["HW"], c["C"], c["ratio"], c["grad_scale"]) if kind == "iw"
             else loss_math.fused_ms(lo, c["HW"], c["grad_scale"]))
        err = np.abs(r["grad_logits"] - ref_g).max()
        assert err <= 1e-4 * np.abs(ref_g).max()
        if kind == "iw":
            gp = loss_tensors[name + "__grad_prob"].astype(np.float64)
            assert np.abs(r["grad_prob"] - gp).max() <= 1e-5 * np.abs(gp).max()


def test_label_argument(loss_kats, loss_tensors):
    """``label=`` changes the counted map but the weights are still gathered by
    argmax(prob) (utils/loss.py:87-96)."""
    c = next(x for x in loss_kats["cases"] if x["name"] == "label_arg")
    lo = torch.from_numpy(loss_tensors["label_arg__logits"])
    lab = torch.from_numpy(loss_tensors["label_arg__label"])
    _, prob = loss_port.prologue(lo, c["HW"])
    loss, hist, _ = loss_port.iw_maxsquare(prob, c["C"], c["ratio"], label=lab, return_aux=True)
    assert hist.tolist() == c["hist"]
    assert abs(loss.item() - c["loss"]) <= 1e-6 * abs(c["loss"])
    p64 = loss_math.softmax64(bilinear.upsample(lo.numpy(), c["HW"]))
    k = loss_math.argmax_of_prob_fp32(bilinear.upsample(lo.numpy(), c["HW"]))
    r = loss_math.iw_from_prob(p64, c["C"], c["ratio"], k=k, hist=loss_math.class_hist_np(lab.numpy(), c["C"]))
    assert abs(r["loss"] - c["loss"]) <= 1e-5 * abs(c["loss"])
    gp = loss_tensors["label_arg__grad_prob"].astype(np.float64)
    assert np.abs(r["grad_prob"] - gp).max() <= 1e-5 * np.abs(gp).max()


def test_histc_is_shifted_bincount():
    lab = synth.random_labels(1, (64, 64), 19, 3)[0]
    lab[0, :5] = 255          # out of range: dropped by histc's [min,max] window
    h = loss_port.class_hist(lab, 19).to(torch.int64).numpy()
    assert h.tolist() == loss_math.class_hist_np(lab[None].numpy(), 19)[0].tolist()


@pytest.mark.parametrize("m", _ECASES, ids=[m["name"] for m in _ECASES])
def test_eval_port_reproduces_reference(m, eval_tensors):
    gt, pr, cm = (eval_tensors[m["name"] + s] for s in ("_gt", "_pr", "_cm"))
    ev = eval_port.EvalPort(m["C"])
    calls = 2 if m["name"].endswith("twice") else 1
    for _ in range(calls):
        ev.add_batch(gt, pr)
    assert np.array_equal(ev.confusion_matrix.astype(np.int64), cm)
    assert ev.confusion_matrix.dtype == np.float64

    def same(a, b):
        a = np.atleast_1d(np.asarray(a, dtype=np.float64))
        b = np.atleast_1d(np.asarray(b, dtype=np.float64))
        return np.array_equal(a, b, equal_nan=True)          # bit-exact, NaN == NaN
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            assert same(ev.Mean_Intersection_over_Union(), m["MIoU"])
            assert same(ev.Pixel_Accuracy(), m["PA"])
            assert same(ev.Mean_Pixel_Accuracy(), m["MPA"])
            assert same(ev.Mean_Precision(), m["PC"])
            assert same(ev.Frequency_Weighted_Intersection_over_Union(), m["FWIoU"])
            if m["C"] == 19:
                assert same(ev.Mean_Intersection_over_Union(out_16_13=True), m["MIoU_16_13"])
                assert same(ev.Frequency_Weighted_Intersection_over_Union(out_16_13=True), m["FWIoU_16_13"])


def test_eval_port_errors():
    ev = eval_port.EvalPort(19)
    with pytest.raises(AssertionError):
        ev.add_batch(np.zeros((1, 4, 4), dtype=np.int64), np.zeros((1, 4, 5), dtype=np.int64))
    with pytest.raises(ValueError):
        ev.add_batch(np.zeros((1, 2, 2), dtype=np.int64), -np.ones((1, 2, 2), dtype=np.int64))
    with pytest.raises(ValueError):
        ev.add_batch(np.full((1, 2, 2), 18, dtype=np.int64), np.full((1, 2, 2), 19, dtype=np.int64))
