"""Source-side fused cross-entropy + evaluation on the GPU (msq_source_ce_fwd / msq_guidance_bwd through
CrossEntropyLoss2d) against the frozen vectors and the oracle.  Bars: confusion matrix, argmax and
n_valid bit-exact; loss <= 1e-5 relative; gradients <= 1e-4 relative."""
import hashlib
import json
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from maxsquareloss_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "source_kats.json")) as _f:
    SOURCE = json.load(_f)["cases"]


@pytest.fixture(scope="module")
def msq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as m
    from maxsquareloss_b200 import _lib, build
    build.build()
    _lib.load()
    return m


def _close(a, b, rtol):
    return math.isnan(a) if math.isnan(b) else abs(a - b) <= rtol * abs(b)


def _grad_close(got, ref, rtol=1e-4):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert (got - ref).abs().max().item() <= rtol * ref.abs().max().item()
    assert (got - ref).norm().item() <= rtol * ref.norm().item()


def _inputs(c):
    lo, y = synth.source_case(c["N"], c["C"], c["hw"], c["HW"], c["seed"], c["scale"], c["label_kind"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"]
    return lo, y


@pytest.mark.parametrize("c", SOURCE, ids=[c["name"] for c in SOURCE])
def test_source_vs_reference_golden(msq, c):
    lo, y = _inputs(c)
    ev = msq.Eval(c["C"])
    crit = msq.CrossEntropyLoss2d(ignore_index=-1, evaluator=ev)
    x = lo.cuda().requires_grad_(True)
    loss = crit(x, y.cuda())
    (c["grad_scale"] * loss).backward()
    assert _close(loss.item(), c["loss"], 1e-5)
    assert int(crit.last_nvalid.item()) == c["nvalid"]
    cm = ev.confusion_matrix.astype(np.int64)
    assert hashlib.sha256(cm.tobytes()).hexdigest() == c["cm_sha256"]            # bit-exact vs the reference's Eval
    with np.errstate(all="ignore"):
        miou = ev.Mean_Intersection_over_Union()
    ref = c["MIoU"]
    if isinstance(ref, list):
        assert list(miou) == ref
    else:
        assert miou == ref or (math.isnan(miou) and math.isnan(ref))
    g = x.grad.double().cpu()
    if c["nvalid"]:
        assert _close(g.abs().sum().item(), c["grad_sum_abs"], 1e-4)
        assert _close(g.norm().item(), c["grad_l2"], 1e-4)
    else:
        assert not g.any()


def test_source_gradients_elementwise_vs_golden(msq):
    t = np.load(os.path.join(GOLDEN, "source_tensors.npz"))
    n = 0
    for c in SOURCE:
        if c["name"] + "__grad" not in t.files or not c["nvalid"]:
            continue
        x = torch.from_numpy(t[c["name"] + "__logits"]).cuda().requires_grad_(True)
        y = torch.from_numpy(t[c["name"] + "__label"].astype(np.int64)).cuda()
        ev = msq.Eval(c["C"])
        loss = msq.CrossEntropyLoss2d()(x, y, evaluator=ev)
        (c["grad_scale"] * loss).backward()
        _grad_close(x.grad, torch.from_numpy(t[c["name"] + "__grad"]))
        assert np.array_equal(ev.confusion_matrix.astype(np.int64), t[c["name"] + "__cm"])
        n += 1
    assert n >= 3


@pytest.mark.parametrize("C,hw,HW,N", [(19, (91, 161), (720, 1280), 2), (16, (96, 161), (760, 1280), 1), (13, (9, 17), (64, 128), 3),
                                       (5, (6, 7), (31, 45), 3), (19, (33, 65), (33, 65), 1), (7, (3, 5), (7, 9), 2),
                                       (21, (10, 12), (40, 150), 1), (32, (8, 8), (64, 64), 1), (2, (4, 4), (17, 300), 1)])
def test_source_vs_oracle(msq, C, hw, HW, N):
    from oracle import eval_port, loss_math
    lo = synth.head_logits(N, C, hw, 91, 2.0)
    y = synth.blocky_labels(N, HW, C, 92, grid=(4, 8))
    y[:, :2, :3] = 255                         # out-of-range labels: ignored by the loss and by Eval
    lo = synth.align_logits_to_labels(lo, torch.where(y == 255, torch.full_like(y, -1), y), boost=2.0)
    m = loss_math.source_ce(lo.numpy(), y.numpy(), 0.3)
    port = eval_port.EvalPort(C)
    port.add_batch(y.numpy(), m["argpred"])
    ev = msq.Eval(C)
    x = lo.cuda().requires_grad_(True)
    crit = msq.CrossEntropyLoss2d(evaluator=ev)
    loss = crit(x, y.cuda())
    (0.3 * loss).backward()
    assert _close(loss.item(), float(m["loss"]), 1e-5)
    assert int(crit.last_nvalid.item()) == m["nvalid"]
    assert np.array_equal(ev.confusion_matrix, port.confusion_matrix)
    _grad_close(x.grad, torch.from_numpy(m["grad_logits"]))


def test_source_equals_torch_cross_entropy_at_full_resolution(msq):
    """pred already at the target's size: the interpolation is the identity and the module is
    nn.CrossEntropyLoss(ignore_index=-1) (the strict call ``self.loss(pred, y)``)."""
    g = torch.Generator().manual_seed(3)
    pred = (torch.randn(2, 19, 40, 72, generator=g) * 3).cuda()
    y = synth.random_labels(2, (40, 72), 19, 4).cuda()
    a = pred.clone().requires_grad_(True)
    ref = F.cross_entropy(a, y, ignore_index=-1)
    ref.backward()
    b = pred.clone().requires_grad_(True)
    ev = msq.Eval(19)
    out = msq.CrossEntropyLoss2d()(b, y, evaluator=ev)
    out.backward()
    assert _close(out.item(), ref.item(), 1e-5)
    _grad_close(b.grad, a.grad)
    ref_ev = msq.Eval(19)
    ref_ev.add_batch(y, pred.argmax(1))
    assert np.array_equal(ev.confusion_matrix, ref_ev.confusion_matrix)


def test_source_errors_and_no_grad(msq):
    x = torch.randn(1, 19, 5, 9, device="cuda")
    y = torch.zeros(1, 33, 65, dtype=torch.int64, device="cuda")
    with torch.no_grad():
        assert torch.isfinite(msq.CrossEntropyLoss2d()(x, y))
    with pytest.raises(RuntimeError):
        msq.CrossEntropyLoss2d()(x.cpu(), y.cpu())
    with pytest.raises(RuntimeError):
        msq.CrossEntropyLoss2d(ignore_index=255)
    with pytest.raises(ValueError):
        msq.CrossEntropyLoss2d(evaluator=msq.Eval(13))(x, y)
    with pytest.raises(RuntimeError):
        msq.CrossEntropyLoss2d()(x, torch.zeros(1, 3, 5, dtype=torch.int64, device="cuda"))     # downsampling
