"""MinEnt losses on the GPU (msq_entropy_fwd / msq_entropy_bwd through softCrossEntropy / IWsoftCrossEntropy)
against the vectors frozen from the reference and against the oracle.  Bars: class histograms bit-exact,
loss <= 1e-5 relative, gradients <= 1e-4 relative."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from maxsquareloss_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "entropy_kats.json")) as _f:
    ENT = json.load(_f)["cases"]


@pytest.fixture(scope="module")
def msq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as m
    from maxsquareloss_b200 import _lib, build
    build.build()
    _lib.load()
    return m


def _crit(msq, iw, C, ratio=0.2):
    return msq.IWsoftCrossEntropy(-1, C, ratio) if iw else msq.softCrossEntropy(-1)


def _grad_close(got, ref, rtol=1e-4):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert (got - ref).abs().max().item() <= rtol * ref.abs().max().item()
    assert (got - ref).norm().item() <= rtol * ref.norm().item()


@pytest.mark.parametrize("c", ENT, ids=[c["name"] for c in ENT])
def test_entropy_vs_reference_golden(msq, c):
    lo = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"], c["quantize"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"]
    x = lo.cuda().requires_grad_(True)
    crit = _crit(msq, c["iw"], c["C"], c["ratio"])
    loss = crit(x, out_size=c["HW"])
    (c["grad_scale"] * loss).backward()
    assert abs(loss.item() - c["loss"]) <= 1e-5 * abs(c["loss"])
    g = x.grad.double().cpu()
    assert abs(g.abs().sum().item() - c["grad_sum_abs"]) <= 1e-4 * c["grad_sum_abs"]
    assert abs(g.norm().item() - c["grad_l2"]) <= 1e-4 * c["grad_l2"]
    if c["iw"]:
        assert crit.last_hist.cpu().tolist() == c["hist"]


def test_entropy_gradients_elementwise_vs_golden(msq):
    t = np.load(os.path.join(GOLDEN, "entropy_tensors.npz"))
    for c in ENT:
        if c["name"] + "__grad" not in t.files:
            continue
        x = torch.from_numpy(t[c["name"] + "__logits"]).cuda().requires_grad_(True)
        (c["grad_scale"] * _crit(msq, c["iw"], c["C"], c["ratio"])(x, out_size=c["HW"])).backward()
        _grad_close(x.grad, torch.from_numpy(t[c["name"] + "__grad"]))


@pytest.mark.parametrize("C,hw,HW,N,scale", [
    (19, (65, 129), (512, 1024), 2, 5.0), (16, (96, 161), (760, 1280), 1, 3.0), (13, (9, 17), (64, 128), 3, 1.0),
    (5, (6, 7), (31, 45), 3, 2.0), (19, (33, 65), (33, 65), 1, 2.0), (7, (3, 5), (7, 9), 2, 1.0),
    (21, (10, 12), (40, 150), 1, 2.0), (32, (8, 8), (64, 64), 1, 2.0), (2, (4, 4), (17, 300), 1, 1.0)])
@pytest.mark.parametrize("iw", [True, False])
def test_entropy_vs_oracle(msq, C, hw, HW, N, scale, iw):
    from oracle import loss_math
    lo = synth.head_logits(N, C, hw, 61, scale)
    r = loss_math.fused_entropy(lo.numpy(), HW, C, iw, 0.2, 0.1)
    x = lo.cuda().requires_grad_(True)
    crit = _crit(msq, iw, C)
    loss = crit(x, out_size=HW)
    (0.1 * loss).backward()
    assert abs(loss.item() - r["loss"]) <= 1e-5 * abs(r["loss"])
    _grad_close(x.grad, torch.from_numpy(r["grad_logits"]))
    if iw:
        assert crit.last_hist.cpu().numpy().tolist() == r["hist"].tolist()


def test_entropy_strict_call_as_the_trainers_make_it(msq):
    """target_loss(pred, softmax(pred)) at full resolution (tools/solve_gta5.py:188-190,199): the model's total
    gradient equals the reference chain's, although `prob` itself receives none from this module."""
    from oracle import loss_port
    lo = synth.head_logits(1, 19, (9, 17), 71, 3.0)
    ref_loss, ref_grad, ref_hist = loss_port.chain_entropy(lo, (64, 128), 19, True, 0.2, 0.1)
    x = lo.cuda().requires_grad_(True)
    pred = F.interpolate(x, size=(64, 128), mode="bilinear", align_corners=True)
    prob = F.softmax(pred, dim=1)
    crit = msq.IWsoftCrossEntropy(-1, 19, 0.2)
    loss = crit(pred, prob)
    (0.1 * loss).backward()
    assert abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
    _grad_close(x.grad, ref_grad)
    assert crit.last_hist.cpu().long().tolist() == ref_hist.tolist()
    with pytest.raises(AssertionError):
        crit(pred, prob[:, :, :32])
    with pytest.raises(RuntimeError):
        crit(x)                                   # fused mode needs out_size
    with pytest.raises(RuntimeError):
        msq.softCrossEntropy()(x.cpu(), out_size=(64, 128))
