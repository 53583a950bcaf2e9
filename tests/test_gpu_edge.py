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


def test_fuzz_one_call_step_against_the_closed_form(msq):
    """Random small geometries through the C ABI's one-call step (two kernels, shared-memory class accumulators, weights derived
    in the backward) against the float64 closed form of the reference chain: class histogram bit-exact, loss and gradient
    within the parity bars; then a NaN logit: NaN loss, clean accumulators, and the next step is right again."""
    from maxsquareloss_b200 import _lib
    from oracle import loss_math
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.RandomState(7)
    for case in range(14):
        n = int(rng.choice([1, 2, 3, 6]))
        C = int(rng.choice([2, 5, 13, 16, 19, 21]))
        h, w = int(rng.randint(2, 12)), int(rng.randint(2, 20))
        H, W = h + int(rng.randint(1, 6 * h)), w + int(rng.randint(1, 7 * w))
        lo = synth.head_logits(n, C, (h, w), 50 + case, float(rng.choice([0.5, 2.0, 6.0])))
        r = loss_math.fused_iw(lo.numpy(), (H, W), C, 0.2, 0.3)
        lay = _lib.state_layout(n, C)
        accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda")
        out = torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda")
        aux = torch.empty(lib.msq_fused_aux_bytes(n, H, W), dtype=torch.uint8, device="cuda")
        x = lo.cuda().contiguous()
        g = torch.full_like(x, float("nan"))

        def step(inp):
            _lib.check(lib.msq_fused_fwd_bwd(_lib.MODE_IW, inp.data_ptr(), n, C, h, w, H, W, 0.2, 0, accum.data_ptr(), out.data_ptr(),
                                             aux.data_ptr(), None, 0.3, g.data_ptr(), None, 0, st))
            torch.cuda.synchronize()
            return out[lay.loss_off:lay.loss_off + 4].view(torch.float32).item()

        what = (case, n, C, (h, w), (H, W))
        loss = step(x)
        hist = out[lay.hist_out_off:lay.hist_out_off + 4 * n * C].view(torch.int32).cpu().numpy().reshape(n, C)
        assert hist.tolist() == np.asarray(r["hist"]).reshape(n, C).tolist(), what
        assert abs(loss - r["loss"]) <= 1e-5 * abs(r["loss"]), what
        ref = torch.from_numpy(r["grad_logits"])
        assert (g.double().cpu() - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-12, what
        bad = x.clone()
        bad[0, C - 1, h - 1, w - 1] = float("nan")
        assert math.isnan(step(bad)), what
        assert not accum.any(), what
        assert step(x) == loss, what
