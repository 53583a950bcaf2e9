"""Multi-level self-produced guidance on the GPU (msq_multi_fwd / msq_guidance_bwd through
MultiLevelTargetLoss) against the vectors frozen from the reference's own train_target source and
against the oracle.  Bars: label_2 / n_valid / class histogram bit-exact, losses <= 1e-5 relative,
gradients <= 1e-4 relative (fp32)."""
import hashlib
import json
import math
import os

import numpy as np
import pytest
import torch

from maxsquareloss_b200 import synth

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "multi_kats.json")) as _f:
    MULTI = json.load(_f)["cases"]


@pytest.fixture(scope="module")
def msq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as m
    from maxsquareloss_b200 import _lib, build
    build.build()
    _lib.load()
    return m


def _module(msq, kind, C, ratio, thr, lam_t, lam_s, **kw):
    tl = msq.IW_MaxSquareloss(-1, C, ratio) if kind == "iw" else msq.MaxSquareloss(-1, C)
    return msq.MultiLevelTargetLoss(tl, threshold=thr, lambda_target=lam_t, lambda_seg=lam_s, return_label=True, **kw)


def _close(a, b, rtol):
    if math.isnan(b):
        return math.isnan(a)
    return abs(a - b) <= rtol * abs(b)


def _grad_close(got, ref, rtol=GRAD_RTOL):
    got, ref = got.double().cpu(), ref.double().cpu()
    assert (got - ref).abs().max().item() <= rtol * ref.abs().max().item()
    assert (got - ref).norm().item() <= rtol * ref.norm().item()


@pytest.mark.parametrize("c", MULTI, ids=[c["name"] for c in MULTI])
def test_multi_vs_reference_golden(msq, c):
    lo1 = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"])
    lo2 = synth.second_head(lo1, c["seed"])
    assert hashlib.sha256(lo2.numpy().tobytes()).hexdigest() == c["input2_sha256"]
    x1, x2 = lo1.cuda().requires_grad_(True), lo2.cuda().requires_grad_(True)
    mod = _module(msq, c["kind"], c["C"], c["ratio"], c["threshold"], c["lambda_target"], c["lambda_seg"])
    lt, lt2 = mod((x1, x2), c["HW"])
    (lt + lt2).backward()
    # integers: bit-exact against the reference's own code
    lab = mod.last_label_2.cpu().numpy()
    assert int(mod.last_nvalid.item()) == c["nvalid"]
    assert np.bincount(lab.reshape(-1) + 1, minlength=c["C"] + 1).tolist() == c["label2_hist"]
    assert hashlib.sha256(lab.tobytes()).hexdigest() == c["label2_sha256"]
    assert _close(lt.item(), c["loss_target"], LOSS_RTOL)
    assert _close(lt2.item(), c["loss_target_2"], LOSS_RTOL)
    g1, g2 = x1.grad.double().cpu(), x2.grad.double().cpu()
    assert _close(g1.abs().sum().item(), c["grad1_sum_abs"], GRAD_RTOL)
    assert _close(g1.norm().item(), c["grad1_l2"], GRAD_RTOL)
    if c["nvalid"]:
        assert _close(g2.abs().sum().item(), c["grad2_sum_abs"], GRAD_RTOL)
        assert _close(g2.norm().item(), c["grad2_l2"], GRAD_RTOL)
    else:
        assert not g2.any()          # NaN loss, all-zero gradient: torch's behaviour, kept


def test_multi_gradients_elementwise_vs_golden(msq):
    t = np.load(os.path.join(GOLDEN, "multi_tensors.npz"))
    n = 0
    for c in MULTI:
        if c["name"] + "__grad2" not in t.files or not c["nvalid"]:
            continue
        x1 = torch.from_numpy(t[c["name"] + "__logits1"]).cuda().requires_grad_(True)
        x2 = torch.from_numpy(t[c["name"] + "__logits2"]).cuda().requires_grad_(True)
        mod = _module(msq, c["kind"], c["C"], c["ratio"], c["threshold"], c["lambda_target"], c["lambda_seg"])
        lt, lt2 = mod((x1, x2), c["HW"])
        (lt + lt2).backward()
        assert np.array_equal(mod.last_label_2.cpu().numpy(), t[c["name"] + "__label2"].astype(np.int64))
        _grad_close(x1.grad, torch.from_numpy(t[c["name"] + "__grad1"]))
        _grad_close(x2.grad, torch.from_numpy(t[c["name"] + "__grad2"]))
        n += 1
    assert n >= 2


@pytest.mark.parametrize("C,hw,HW,N,scale,thr", [
    (19, (65, 129), (512, 1024), 2, 5.0, 0.95), (16, (96, 161), (760, 1280), 1, 4.0, 0.98),
    (13, (9, 17), (64, 128), 3, 4.0, 0.9), (5, (6, 7), (31, 45), 3, 3.0, 0.8), (19, (33, 65), (33, 65), 1, 6.0, 0.95),
    (7, (3, 5), (7, 9), 2, 3.0, 0.5), (21, (10, 12), (40, 150), 1, 6.0, 0.9), (32, (8, 8), (64, 64), 1, 8.0, 0.9),
    (2, (4, 4), (17, 300), 1, 2.0, 0.7), (8, (12, 20), (12, 131), 2, 4.0, 0.9)])
@pytest.mark.parametrize("kind", ["iw", "ms"])
def test_multi_vs_oracle(msq, C, hw, HW, N, scale, thr, kind):
    from oracle import loss_math
    lo1 = synth.head_logits(N, C, hw, 77, scale)
    lo2 = synth.second_head(lo1, 77)
    r1 = (loss_math.fused_iw(lo1.numpy(), HW, C, 0.2, 0.09) if kind == "iw" else loss_math.fused_ms(lo1.numpy(), HW, 0.09))
    r2 = loss_math.guidance(lo1.numpy(), lo2.numpy(), HW, thr, 0.009)
    x1, x2 = lo1.cuda().requires_grad_(True), lo2.cuda().requires_grad_(True)
    mod = _module(msq, kind, C, 0.2, thr, 0.09, 0.1)
    lt, lt2 = mod((x1, x2), HW)
    (lt + lt2).backward()
    assert np.array_equal(mod.last_label_2.cpu().numpy(), r2["label_2"])
    assert int(mod.last_nvalid.item()) == r2["nvalid"]
    assert _close(lt.item(), 0.09 * r1["loss"], LOSS_RTOL)
    assert _close(lt2.item(), 0.009 * float(r2["loss2"]), LOSS_RTOL)
    _grad_close(x1.grad, torch.from_numpy(r1["grad_logits"]))
    if r2["nvalid"]:
        _grad_close(x2.grad, torch.from_numpy(r2["grad_logits2"]))
    if kind == "iw":
        assert mod.target_loss.last_hist.cpu().numpy().tolist() == r1["hist"].tolist()


def test_multi_vs_torch_cuda_eager_chain(msq):
    """The reference's op chain executed by torch eager on the SAME GPU: label_2 bit-exact."""
    from oracle import loss_port
    for seed, scale, thr in [(0, 5.0, 0.95), (1, 3.0, 0.9), (2, 8.0, 0.98), (3, 1.0, 0.3)]:
        lo1 = synth.head_logits(1, 19, (65, 129), seed, scale)
        lo2 = synth.second_head(lo1, seed)
        r = loss_port.chain_multi(lo1.cuda(), lo2.cuda(), (512, 1024), 19, "iw", 0.2, thr, 0.1, 0.1)
        x1, x2 = lo1.cuda().requires_grad_(True), lo2.cuda().requires_grad_(True)
        mod = _module(msq, "iw", 19, 0.2, thr, 0.1, 0.1)
        lt, lt2 = mod((x1, x2), (512, 1024))
        (lt + lt2).backward()
        assert torch.equal(mod.last_label_2, r["label_2"]), (seed, scale)
        assert _close(lt.item(), r["loss_target"].item(), LOSS_RTOL)
        assert _close(lt2.item(), r["loss_target_2"].item(), LOSS_RTOL)
        _grad_close(x1.grad, r["grad1"])
        _grad_close(x2.grad, r["grad2"])


def test_multi_head1_equals_single_head_kernels(msq):
    """Head 1 of the multi kernel is the single-head fused loss: identical loss bits and histogram,
    gradient within fp32 atomics noise."""
    lo1 = synth.head_logits(2, 19, (65, 129), 5, 5.0)
    lo2 = synth.second_head(lo1, 5)
    xs = lo1.cuda().requires_grad_(True)
    crit = msq.IW_MaxSquareloss(-1, 19, 0.2)
    ls = crit(xs, out_size=(512, 1024))
    ls.backward()
    x1, x2 = lo1.cuda().requires_grad_(True), lo2.cuda().requires_grad_(True)
    mod = _module(msq, "iw", 19, 0.2, 0.95, 1.0, 0.1)
    lt, lt2 = mod((x1, x2), (512, 1024))
    lt.backward()                      # only head 1 contributes
    assert lt.item() == ls.item()
    assert torch.equal(mod.target_loss.last_hist, crit.last_hist)
    _grad_close(x1.grad, xs.grad, 1e-5)
    assert x2.grad is None or not x2.grad.any()


def test_multi_sharded_mean_composes(msq):
    """CE is a mean over the valid pixels of the WHOLE batch: {sum, count} of per-shard calls add up
    to the full-batch call (what the all-reduce in _MultiLoss does across ranks)."""
    lo1 = synth.head_logits(4, 19, (33, 65), 9, 5.0)
    lo2 = synth.second_head(lo1, 9)
    full = _module(msq, "iw", 19, 0.2, 0.95, 1.0, 1.0, group=False)
    full.target_loss.global_batch = 4
    _, l2 = full((lo1.cuda(), lo2.cuda()), (257, 513))
    ce, nv = 0.0, 0
    for lo, hi in ((0, 1), (1, 4)):
        part = _module(msq, "iw", 19, 0.2, 0.95, 1.0, 1.0, group=False)
        part.target_loss.global_batch = 4
        part((lo1[lo:hi].cuda(), lo2[lo:hi].cuda()), (257, 513))
        ce += part.last_ce_sum.item()
        nv += int(part.last_nvalid.item())
    assert nv == int(full.last_nvalid.item())
    assert abs(ce / nv - l2.item()) <= 1e-6 * abs(l2.item())
    assert abs(ce - full.last_ce_sum.item()) <= 1e-9 * abs(ce)      # fixed-point sums: order-independent


def test_multi_errors(msq):
    mod = _module(msq, "iw", 19, 0.2, 0.95, 0.1, 0.1)
    a = torch.zeros(1, 19, 4, 4, device="cuda")
    with pytest.raises(RuntimeError):
        mod(a, (8, 8))                                   # needs both heads
    with pytest.raises(RuntimeError):
        mod((a, torch.zeros(1, 19, 4, 5, device="cuda")), (8, 8))
    with pytest.raises(RuntimeError):
        mod((a.cpu(), a.cpu()), (8, 8))                  # no CPU fallback
    with pytest.raises(ValueError):
        mod((torch.zeros(1, 13, 4, 4, device="cuda"),) * 2, (8, 8))
    with pytest.raises(TypeError):
        msq.MultiLevelTargetLoss(torch.nn.CrossEntropyLoss())
