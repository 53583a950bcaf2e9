"""Edge cases of the fused path the reference's callers can produce: non-finite logits, saturated softmax, many tiny
images per CTA, one class, identity geometry in one axis, non-contiguous inputs, a zero upstream gradient."""
import math

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
    build.build()
    _lib.load()
    return m


def _grad_close(got, ref, rtol=1e-4):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert (got - ref).abs().max().item() <= rtol * ref.abs().max().item()


def test_nan_logits_give_nan_loss_and_the_state_recovers(msq):
    """The trainer raises on a NaN loss (tools/train_source.py:277-278); the accumulators must be clean afterwards."""
    lo = synth.head_logits(1, 19, (9, 17), 1, 2.0).cuda()
    bad = lo.clone()
    bad[0, 3, 4, 5] = float("nan")
    for crit in (msq.IW_MaxSquareloss(-1, 19, 0.2), msq.MaxSquareloss(-1, 19)):
        good = crit(lo, out_size=(64, 128)).item()
        assert math.isnan(crit(bad, out_size=(64, 128)).item())
        assert crit(lo, out_size=(64, 128)).item() == good
    a = msq.CrossEntropyLoss2d()(lo, torch.zeros(1, 64, 128, dtype=torch.int64, device="cuda")).item()
    assert math.isnan(msq.CrossEntropyLoss2d()(bad, torch.zeros(1, 64, 128, dtype=torch.int64, device="cuda")).item())
    assert msq.CrossEntropyLoss2d()(lo, torch.zeros(1, 64, 128, dtype=torch.int64, device="cuda")).item() == a


def test_saturated_logits(msq):
    """|logits| ~ 1e3: softmax is one-hot, exp underflows to 0 for the losers; loss -> -1/C * sum w, gradient -> 0."""
    from oracle import loss_math
    lo = synth.head_logits(1, 13, (9, 17), 2, 300.0)
    r = loss_math.fused_iw(lo.numpy(), (64, 128), 13, 0.2, 1.0)
    x = lo.cuda().requires_grad_(True)
    crit = msq.IW_MaxSquareloss(-1, 13, 0.2)
    loss = crit(x, out_size=(64, 128))
    loss.backward()
    assert abs(loss.item() - r["loss"]) <= 1e-5 * abs(r["loss"])
    assert crit.last_hist.cpu().numpy().tolist() == r["hist"].tolist()
    assert torch.isfinite(x.grad).all()
    ref = torch.from_numpy(r["grad_logits"])
    assert (x.grad.double().cpu() - ref).abs().max().item() <= 1e-4 * max(ref.abs().max().item(), 1e-12) + 1e-12


@pytest.mark.parametrize("N,C,hw,HW", [(64, 19, (3, 3), (8, 8)), (33, 13, (2, 5), (5, 9)), (3, 1, (4, 6), (9, 20)),
                                       (2, 19, (16, 20), (16, 100)), (2, 16, (7, 50), (40, 50)), (1, 19, (2, 2), (300, 2))])
def test_many_tiny_images_and_degenerate_geometry(msq, N, C, hw, HW):
    from oracle import loss_math
    lo = synth.head_logits(N, C, hw, 4, 2.0)
    for kind in ("iw", "ms"):
        r = loss_math.fused_iw(lo.numpy(), HW, C, 0.2, 0.3) if kind == "iw" else loss_math.fused_ms(lo.numpy(), HW, 0.3)
        x = lo.cuda().requires_grad_(True)
        crit = msq.IW_MaxSquareloss(-1, C, 0.2) if kind == "iw" else msq.MaxSquareloss(-1, C)
        loss = crit(x, out_size=HW)
        (0.3 * loss).backward()
        assert abs(loss.item() - r["loss"]) <= 1e-5 * abs(r["loss"])
        _grad_close(x.grad, torch.from_numpy(r["grad_logits"]))
        if kind == "iw":
            assert crit.last_hist.cpu().numpy().tolist() == r["hist"].tolist()


def test_noncontiguous_logits_and_zero_upstream_gradient(msq):
    from oracle import loss_math
    base = synth.head_logits(2, 19, (17, 9), 6, 3.0)              # stored (N,C,w,h)
    lo = base.permute(0, 1, 3, 2)                                   # a transposed, non-contiguous view (N,C,9,17)
    r = loss_math.fused_iw(np.ascontiguousarray(lo.numpy()), (64, 128), 19, 0.2, 1.0)
    x = base.cuda().requires_grad_(True)
    crit = msq.IW_MaxSquareloss(-1, 19, 0.2)
    loss = crit(x.permute(0, 1, 3, 2), out_size=(64, 128))
    loss.backward()
    assert abs(loss.item() - r["loss"]) <= 1e-5 * abs(r["loss"])
    _grad_close(x.grad.permute(0, 1, 3, 2), torch.from_numpy(r["grad_logits"]))
    y = base.cuda().requires_grad_(True)
    (0.0 * crit(y.permute(0, 1, 3, 2), out_size=(64, 128))).backward()
    assert not y.grad.any()
