"""Strict MinEnt kernels (msq_softce_fwd / msq_softce_bwd through softCrossEntropy / IWsoftCrossEntropy called with a
target tensor) against vectors frozen from the reference's own classes and against the oracle port.  Bars: class
histograms bit-exact, loss <= 1e-5 relative, gradients (both arguments) <= 1e-4 relative."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import loss_port

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "softce_kats.json")) as _f:
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


def _close(got, ref, rtol=1e-4):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert (got - ref).abs().max().item() <= rtol * ref.abs().max().item()
    assert (got - ref).norm().item() <= rtol * ref.norm().item()


@pytest.mark.parametrize("c", CASES, ids=[c["name"] for c in CASES])
def test_strict_minent_vs_reference_golden(msq, c):
    t = np.load(os.path.join(GOLDEN, "softce_tensors.npz"))
    x = torch.from_numpy(t[c["name"] + "__inputs"]).cuda().requires_grad_(True)
    if c["target"] == "self":                        # the trainers' call: target = softmax(inputs), attached
        tt = F.softmax(x, dim=1)
        tt.retain_grad()
    else:
        tt = torch.from_numpy(t[c["name"] + "__target"]).cuda().requires_grad_(True)
    crit = msq.IWsoftCrossEntropy(-1, c["C"], c["ratio"]) if c["iw"] else msq.softCrossEntropy(-1)
    loss = crit(x, tt)
    (c["grad_scale"] * loss).backward()
    assert abs(loss.item() - c["loss"]) <= 1e-5 * abs(c["loss"])
    if c["iw"]:
        assert crit.last_hist.cpu().tolist() == c["hist"]
    _close(x.grad, torch.from_numpy(t[c["name"] + "__grad_inputs"]))
    _close(tt.grad, torch.from_numpy(t[c["name"] + "__grad_target"]))


@pytest.mark.parametrize("iw", [False, True])
@pytest.mark.parametrize("C,N,HW", [(19, 2, (64, 128)), (16, 1, (33, 65)), (5, 3, (8, 7)), (32, 1, (16, 16))])
def test_strict_minent_vs_port(msq, iw, C, N, HW):
    """Sizes with even / odd H*W (64-bit and scalar load paths), specialised and generic class counts, N > 1 for the IW
    loss (defined as the mean over images of the N = 1 loss), a detached target (no gradient for it)."""
    g = torch.Generator().manual_seed(7 * C + N)
    z = torch.randn(N, C, *HW, generator=g) * 3
    tgt = F.softmax(torch.randn(N, C, *HW, generator=g), 1)
    tgt[:, 0, 0, :3] = -1.0                           # a few masked elements
    xr = z.clone().requires_grad_(True)
    if iw:
        ref, hist = loss_port.iw_soft_cross_entropy(xr, tgt, C, 0.2, return_aux=True)
    else:
        ref = loss_port.soft_cross_entropy(xr, tgt)
    (0.3 * ref).backward()
    x = z.cuda().requires_grad_(True)
    crit = msq.IWsoftCrossEntropy(-1, C, 0.2) if iw else msq.softCrossEntropy(-1)
    loss = crit(x, tgt.cuda())
    (0.3 * loss).backward()
    assert abs(loss.item() - ref.item()) <= 1e-5 * abs(ref.item())
    _close(x.grad, xr.grad)
    if iw:
        assert crit.last_hist.cpu().long().tolist() == hist.tolist()


def test_strict_minent_differs_from_the_fused_shortcut_when_the_target_is_not_softmax(msq):
    """Round 1 ignored `target` (it recomputed softmax(inputs)); a uniform target must give the uniform-target loss."""
    C, HW = 19, (16, 32)
    z = torch.randn(1, C, *HW, generator=torch.Generator().manual_seed(3)) * 4
    uniform = torch.full_like(z, 1.0 / C)
    ref = loss_port.soft_cross_entropy(z, uniform)
    fused = msq.softCrossEntropy(-1)(z.cuda(), out_size=HW)            # MinEnt proper: target = softmax
    strict = msq.softCrossEntropy(-1)(z.cuda(), uniform.cuda())
    assert abs(strict.item() - ref.item()) <= 1e-5 * abs(ref.item())
    assert abs(strict.item() - fused.item()) > 0.1 * abs(ref.item())


def test_strict_minent_errors(msq):
    z = torch.randn(1, 19, 8, 8).cuda()
    with pytest.raises(AssertionError):
        msq.softCrossEntropy(-1)(z, torch.randn(1, 19, 8, 4).cuda())          # utils/loss.py:29
    with pytest.raises(RuntimeError):
        msq.softCrossEntropy(-1)(z, torch.randn(1, 19, 8, 8))                 # CPU target: no fallback
