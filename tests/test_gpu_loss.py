"""CUDA loss kernels (strict drop-in and fused) against the oracle and the vectors frozen from
the reference.  Tolerances are the north star's: class histograms bit-exact, loss <= 1e-5
relative, gradients <= 1e-4 relative (fp32)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from maxsquareloss_b200 import synth

pytestmark = pytest.mark.gpu
LOSS_RTOL = 1e-5
GRAD_RTOL = 1e-4

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
with open(os.path.join(GOLDEN, "loss_kats.json")) as _f:
    _CASES = [c for c in json.load(_f)["cases"] if c["kind"] in ("iw", "ms")]


@pytest.fixture(scope="module")
def msq():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as m
    from maxsquareloss_b200 import _lib, build
    build.build()        # no-op when the in-tree library is current
    _lib.load()
    return m


def _crit(msq, kind, C, ratio=0.2):
    return msq.IW_MaxSquareloss(-1, C, ratio) if kind == "iw" else msq.MaxSquareloss(-1, C)


def _grad_close(got, ref, rtol=GRAD_RTOL):
    got, ref = got.double().cpu(), ref.double().cpu()
    scale = ref.abs().max().item()
    assert (got - ref).abs().max().item() <= rtol * scale
    assert (got - ref).norm().item() <= rtol * ref.norm().item()


# ------------------------------------------------------------------ fused path vs the frozen reference
@pytest.mark.parametrize("c", _CASES, ids=[c["name"] for c in _CASES])
def test_fused_vs_reference_golden(msq, c):
    lo = synth.head_logits(c["N"], c["C"], c["hw"], c["seed"], c["scale"], c["class_bias"], c["quantize"])
    assert hashlib.sha256(lo.numpy().tobytes()).hexdigest() == c["input_sha256"]
    x = lo.cuda().requires_grad_(True)
    crit = _crit(msq, c["kind"], c["C"], c["ratio"])
    loss = crit(x, out_size=c["HW"])
    (c["grad_scale"] * loss).backward()
    assert abs(loss.item() - c["loss"]) <= LOSS_RTOL * abs(c["loss"])
    g = x.grad.double().cpu()
    assert abs(g.abs().sum().item() - c["grad_sum_abs"]) <= GRAD_RTOL * c["grad_sum_abs"]
    assert abs(g.norm().item() - c["grad_l2"]) <= GRAD_RTOL * c["grad_l2"]
    if c["kind"] == "iw":
        # histogram of the REFERENCE (CPU torch) on the same input, bit-exact -- including
        # the cases with exact ties between interpolated logits (first index wins)
        assert crit.last_hist.cpu().tolist() == c["hist"]


def test_fused_gradients_elementwise_vs_golden(msq, loss_tensors):
    for name, kind in (("KAT5_iw_c13_tiny", "iw"), ("ms_c13_tiny", "ms")):
        c = next(x for x in _CASES if x["name"] == name)
        x = torch.from_numpy(loss_tensors[name + "__logits"]).cuda().requires_grad_(True)
        crit = _crit(msq, kind, c["C"], c["ratio"])
        (c["grad_scale"] * crit(x, out_size=c["HW"])).backward()
        _grad_close(x.grad, torch.from_numpy(loss_tensors[name + "__grad_logits"]))


# ------------------------------------------------------------------ fused path vs the oracle, more geometry
@pytest.mark.parametrize("C,hw,HW,N,scale", [
    (19, (65, 129), (512, 1024), 2, 5.0), (16, (96, 161), (760, 1280), 1, 3.0), (19, (81, 161), (640, 1280), 1, 5.0),
    (13, (9, 17), (64, 128), 3, 1.0), (5, (6, 7), (31, 45), 3, 2.0), (19, (33, 65), (33, 65), 1, 2.0),
    (7, (3, 5), (7, 9), 2, 1.0), (21, (10, 12), (40, 150), 1, 2.0), (32, (8, 8), (64, 64), 1, 2.0),
    (2, (4, 4), (17, 300), 1, 1.0), (19, (1, 1), (5, 7), 1, 1.0), (8, (12, 20), (12, 131), 2, 3.0)])
@pytest.mark.parametrize("kind", ["iw", "ms"])
def test_fused_vs_oracle(msq, C, hw, HW, N, scale, kind):
    from oracle import loss_math
    lo = synth.head_logits(N, C, hw, 42, scale)
    r = (loss_math.fused_iw(lo.numpy(), HW, C, 0.2, 0.1) if kind == "iw" else loss_math.fused_ms(lo.numpy(), HW, 0.1))
    x = lo.cuda().requires_grad_(True)
    crit = _crit(msq, kind, C)
    loss = crit(x, out_size=HW)
    (0.1 * loss).backward()
    assert abs(loss.item() - r["loss"]) <= LOSS_RTOL * abs(r["loss"])
    _grad_close(x.grad, torch.from_numpy(r["grad_logits"]))
    if kind == "iw":
        # oracle argmax on the bit-exact interpolated logits; tiny outputs use the CUDA/large-tensor
        # arithmetic (oracle/bilinear.py), which is what the kernel implements
        assert crit.last_hist.cpu().numpy().tolist() == r["hist"].tolist()
        w = loss_math.weights_fp32(r["hist"], 0.2)
        assert np.abs(crit.last_weights.cpu().numpy() - w).max() <= 2e-6 * np.abs(w).max()   # fp32 pow differs by a few ulp between libraries


def test_fused_vs_torch_cuda_eager_chain(msq):
    """The reference's op chain executed by torch eager on the SAME GPU: histogram bit-exact,
    including small-magnitude logits where softmax outputs tie although logits do not."""
    from oracle import loss_port
    for seed, scale in [(0, 1.0), (1, 5.0), (2, 0.1), (3, 0.02), (4, 0.5)]:
        lo = synth.head_logits(1, 19, (65, 129), seed, scale).cuda()
        ref_loss, ref_grad, ref_hist = loss_port.chain_iw_maxsquare(lo, (512, 1024), 19, 0.2, 0.1)
        x = lo.clone().requires_grad_(True)
        crit = msq.IW_MaxSquareloss(-1, 19, 0.2)
        loss = crit(x, out_size=(512, 1024))
        (0.1 * loss).backward()
        assert crit.last_hist.cpu().long().tolist() == ref_hist.cpu().tolist(), (seed, scale)
        assert abs(loss.item() - ref_loss.item()) <= LOSS_RTOL * abs(ref_loss.item())
        _grad_close(x.grad, ref_grad)


def test_fused_near_tie_accounting_vs_cpu(msq):
    """Against the CPU reference chain at real size with tiny logits: every histogram difference
    must be explained by pixels whose fp32 softmax outputs tie (counted by the oracle)."""
    from oracle import bilinear, loss_math
    lo = synth.head_logits(1, 19, (65, 129), 2, 0.1)
    z = bilinear.upsample(lo.numpy(), (512, 1024))
    k_ref = loss_math.argmax_of_prob_fp32(z)                       # what the CPU reference counts
    k_logit = z.argmax(axis=1)
    suspects = int((k_ref != k_logit).sum())                        # p-ties resolved to a lower index
    crit = msq.IW_MaxSquareloss(-1, 19, 0.2)
    crit(lo.cuda(), out_size=(512, 1024))
    hist_ref = loss_math.class_hist_np(k_ref, 19)
    diff = np.abs(crit.last_hist.cpu().numpy().astype(np.int64) - hist_ref).sum()
    assert diff <= 2 * suspects + 2 * int(loss_math.near_tie_pixels(z, ulps=2).sum())


def test_fused_label_argument_and_batches(msq, loss_kats, loss_tensors):
    c = next(x for x in loss_kats["cases"] if x["name"] == "label_arg")
    x = torch.from_numpy(loss_tensors["label_arg__logits"]).cuda().requires_grad_(True)
    lab = torch.from_numpy(loss_tensors["label_arg__label"]).cuda()
    crit = msq.IW_MaxSquareloss(-1, c["C"], c["ratio"])
    loss = crit(x, label=lab, out_size=c["HW"])
    loss.backward()
    assert crit.last_hist.cpu().tolist() == c["hist"]
    assert abs(loss.item() - c["loss"]) <= LOSS_RTOL * abs(c["loss"])
    _grad_close(x.grad, torch.from_numpy(loss_tensors["label_arg__grad_logits"]))


def test_fused_batch_is_mean_of_images_and_deterministic(msq):
    lo = synth.head_logits(8, 19, (65, 129), 7, 4.0, class_bias=True).cuda()
    crit = msq.IW_MaxSquareloss(-1, 19, 0.2)
    whole = crit(lo, out_size=(512, 1024))
    hist = crit.last_hist.clone()
    singles = []
    for i in range(8):
        singles.append(crit(lo[i:i + 1], out_size=(512, 1024)).item())
        assert torch.equal(crit.last_hist[0], hist[i])
    assert abs(whole.item() - np.mean(singles)) <= 1e-6 * abs(whole.item())
    again = crit(lo, out_size=(512, 1024))
    assert again.item() == whole.item()                 # integer accumulation: bit-reproducible
    # sharded normalisation: partial losses with global_batch sum to the whole
    crit.global_batch = 8
    parts = [crit(lo[a:a + 2], out_size=(512, 1024)).item() for a in range(0, 8, 2)]
    assert abs(sum(parts) - whole.item()) <= 1e-6 * abs(whole.item())


def test_fused_backward_with_and_without_stats_cache(msq):
    """The backward either recomputes max/argmax/normaliser from the logits or reads the 16 B/pixel
    cache the forward wrote; both must agree (and both are within tolerance of the oracle)."""
    from maxsquareloss_b200 import loss as L
    from oracle import loss_math
    for kind, C, hw, HW in (("iw", 19, (65, 129), (512, 1024)), ("ms", 19, (65, 129), (512, 1024)),
                            ("iw", 13, (9, 17), (64, 128)), ("iw", 5, (6, 7), (31, 45))):
        lo = synth.head_logits(2, C, hw, 77, 4.0)
        ref = (loss_math.fused_iw(lo.numpy(), HW, C, 0.2, 0.1) if kind == "iw" else loss_math.fused_ms(lo.numpy(), HW, 0.1))
        grads = []
        try:
            for use in (True, False):
                L.USE_STATS_CACHE = use
                x = lo.cuda().requires_grad_(True)
                crit = _crit(msq, kind, C)
                (0.1 * crit(x, out_size=HW)).backward()
                _grad_close(x.grad, torch.from_numpy(ref["grad_logits"]))
                grads.append(x.grad.clone())
        finally:
            L.USE_STATS_CACHE = True
        _grad_close(grads[0], grads[1], rtol=1e-5)
    # retain_graph: the second backward cannot reuse the pre-zeroed buffer
    x = synth.head_logits(1, 13, (9, 17), 5, 2.0).cuda().requires_grad_(True)
    loss = msq.IW_MaxSquareloss(-1, 13, 0.2)(x, out_size=(64, 128))
    loss.backward(retain_graph=True)
    g1 = x.grad.clone()
    x.grad = None
    loss.backward()
    _grad_close(x.grad, g1, rtol=1e-6)


def test_fused_tuning_knob_does_not_change_results(msq):
    from maxsquareloss_b200 import _lib
    lo = synth.head_logits(1, 19, (65, 129), 3, 3.0).cuda()
    crit = msq.IW_MaxSquareloss(-1, 19, 0.2)
    base = None
    try:
        for rows in (0, 3, 8, 17, 64):
            _lib.tune("fused_rows", rows)
            x = lo.clone().requires_grad_(True)
            loss = crit(x, out_size=(512, 1024))
            loss.backward()
            cur = (loss.item(), crit.last_hist.clone(), x.grad.clone())
            if base is None:
                base = cur
            else:
                assert torch.equal(cur[1], base[1])
                assert abs(cur[0] - base[0]) <= 1e-6 * abs(base[0])
                _grad_close(cur[2], base[2], rtol=1e-5)
    finally:
        _lib.tune("fused_rows", 0)


# ------------------------------------------------------------------ strict drop-in
@pytest.mark.parametrize("C,hw,HW,N", [(19, (65, 129), (512, 1024), 2), (16, (96, 161), (760, 1280), 1),
                                       (13, (9, 17), (64, 128), 1), (5, (6, 7), (31, 45), 3), (21, (5, 5), (33, 35), 2),
                                       (32, (4, 4), (16, 16), 1)])
@pytest.mark.parametrize("kind", ["iw", "ms"])
def test_strict_vs_port(msq, C, hw, HW, N, kind):
    """forward(pred, prob) exactly as tools/solve_gta5.py:199 calls it."""
    from oracle import loss_port
    lo = synth.head_logits(N, C, hw, 1, 3.0)
    pred, prob = loss_port.prologue(lo, HW)
    p = prob.clone().requires_grad_(True)
    if kind == "iw":
        ref, ref_hist, _ = loss_port.iw_maxsquare(p, C, 0.2, return_aux=True)
    else:
        ref, ref_hist = loss_port.maxsquare(p), None
    (0.09 * ref).backward()
    pg = prob.cuda().requires_grad_(True)
    crit = _crit(msq, kind, C)
    loss = crit(pred.cuda(), pg)
    (0.09 * loss).backward()
    assert loss.dim() == 0 and loss.is_cuda
    assert abs(loss.item() - ref.item()) <= LOSS_RTOL * abs(ref.item())
    _grad_close(pg.grad, p.grad)
    if kind == "iw":
        assert torch.equal(crit.last_hist.cpu().long(), ref_hist)       # unconditional: argmax of the given prob


def test_strict_golden_grad_prob(msq, loss_tensors, loss_kats):
    from oracle import loss_port
    for name in ("KAT5_iw_c13_tiny", "label_arg"):
        c = next(x for x in loss_kats["cases"] if x["name"] == name)
        lo = torch.from_numpy(loss_tensors[name + "__logits"])
        _, prob = loss_port.prologue(lo, c["HW"])
        pg = prob.cuda().requires_grad_(True)
        lab = torch.from_numpy(loss_tensors["label_arg__label"]).cuda() if name == "label_arg" else None
        crit = msq.IW_MaxSquareloss(-1, c["C"], c["ratio"])
        loss = crit(None, pg, lab)
        loss.backward()
        assert abs(loss.item() - c["loss"]) <= LOSS_RTOL * abs(c["loss"])
        assert crit.last_hist.cpu().tolist() == c["hist"]
        _grad_close(pg.grad, torch.from_numpy(loss_tensors[name + "__grad_prob"]))


def test_strict_ignore_mask_and_noncontiguous(msq):
    """prob entries equal to ignore_index are masked exactly as utils/loss.py:85-86,117 do."""
    from oracle import loss_port
    g = torch.Generator().manual_seed(5)
    prob = torch.softmax(torch.randn(2, 19, 24, 40, generator=g), 1)
    prob[0, :, 3, 5] = -1.0                   # whole pixel ignored (its max is -1)
    prob[1, 4, 7, 7] = -1.0                   # one element ignored
    for kind in ("iw", "ms"):
        p = prob.clone().requires_grad_(True)
        ref = loss_port.iw_maxsquare(p, 19, 0.2) if kind == "iw" else loss_port.maxsquare(p)
        ref.backward()
        base = prob.permute(0, 2, 3, 1).contiguous().cuda()
        pg = base.permute(0, 3, 1, 2).requires_grad_(True)      # NCHW view of NHWC memory
        crit = _crit(msq, kind, 19)
        loss = crit(None, pg)
        loss.backward()
        assert abs(loss.item() - ref.item()) <= LOSS_RTOL * abs(ref.item())
        _grad_close(pg.grad, p.grad)


def test_strict_nan_propagates_and_state_recovers(msq):
    prob = torch.softmax(torch.randn(1, 19, 16, 16), 1).cuda()
    bad = prob.clone()
    bad[0, 3, 2, 2] = float("nan")
    crit = msq.MaxSquareloss(-1, 19)
    assert np.isnan(crit(None, bad).item())            # trainer raises on NaN loss (train_source.py:277)
    ok = crit(None, prob).item()
    assert np.isfinite(ok) and ok == crit(None, prob).item()


def test_api_contract(msq):
    crit = msq.IW_MaxSquareloss(-1, 19, 0.2).to("cuda")             # .to(device) as in solve_gta5.py:165
    lo = synth.head_logits(1, 19, (9, 17), 1).cuda().requires_grad_(True)
    pred = F.interpolate(lo, size=(64, 128), mode="bilinear", align_corners=True)
    prob = F.softmax(pred, dim=1)
    loss = 0.1 * crit(pred, prob)                                    # loss * float, then autograd through
    loss.backward()                                                  # torch's own softmax / interpolate
    assert lo.grad is not None and torch.isfinite(lo.grad).all()
    lo2 = lo.detach().clone().requires_grad_(True)
    loss2 = 0.1 * crit(lo2, out_size=(64, 128))
    loss2.backward()
    assert abs(loss2.item() - loss.item()) <= LOSS_RTOL * abs(loss.item())
    _grad_close(lo2.grad, lo.grad)                                   # fused == strict + torch autograd
    with pytest.raises(ValueError):
        crit(None, torch.softmax(torch.randn(1, 16, 8, 8), 1).cuda())
    with pytest.raises(RuntimeError):
        crit(None, prob.double())
    with pytest.raises(RuntimeError):
        crit(lo, out_size=(4, 4))                                    # downsampling is not supported


def test_host_pipeline_matches_oracle(msq):
    """C-ABI host-buffer pipeline (msq_pipe_*): pinned host logits in, loss / hist / dL/dlogits out."""
    from oracle import loss_port
    C, hw, HW = synth.SHAPES["tiny13"]
    pipe = msq.HostPipeline("iw", 2, C, hw, HW, ratio=0.2, depth=2)
    refs, outs = [], []
    for seed in range(5):                     # more submissions than slots: slots are recycled
        lo = synth.head_logits(2, C, hw, 50 + seed, 2.0)
        refs.append(loss_port.chain_iw_maxsquare(lo, HW, C, 0.2, 0.1))
        bufs = (lo.pin_memory(), torch.empty(()).pin_memory(), torch.empty_like(lo).pin_memory(),
                torch.empty(2, C, dtype=torch.int32).pin_memory())
        slot = pipe.submit(bufs[0], bufs[1], bufs[2], bufs[3], grad_scale=0.1)
        outs.append((slot, bufs))
        if seed >= 1:                          # read the previous submission while this one runs
            pslot, pb = outs[seed - 1]
            pipe.wait(pslot)
            rl, rg, rh = refs[seed - 1]
            assert abs(pb[1].item() - rl.item()) <= LOSS_RTOL * abs(rl.item())
            assert pb[3].long().tolist() == rh.tolist()
            _grad_close(pb[2], rg)
    pipe.drain()
    rl, rg, rh = refs[-1]
    assert abs(outs[-1][1][1].item() - rl.item()) <= LOSS_RTOL * abs(rl.item())
    _grad_close(outs[-1][1][2], rg)
    pipe.close()
    ms = msq.HostPipeline("maxsquare", 1, C, hw, HW)
    lo = synth.head_logits(1, C, hw, 9, 2.0)
    rl, rg = loss_port.chain_maxsquare(lo, HW, 1.0)
    loss, grad = torch.empty(()).pin_memory(), torch.empty_like(lo).pin_memory()
    ms.wait(ms.submit(lo.pin_memory(), loss, grad))
    assert abs(loss.item() - rl.item()) <= LOSS_RTOL * abs(rl.item())
    _grad_close(grad, rg)


def test_one_call_step_equals_forward_plus_backward(msq):
    """C ABI msq_fused_fwd_bwd (one library call per step) == the nn.Module's forward + backward."""
    from maxsquareloss_b200 import _lib
    lib = _lib.load()
    for kind, mode in (("iw", _lib.MODE_IW), ("ms", _lib.MODE_MAXSQUARE)):
        lo = synth.head_logits(2, 19, (33, 65), 12, 4.0).cuda()
        x = lo.clone().requires_grad_(True)
        crit = _crit(msq, kind, 19)
        loss = crit(x, out_size=(257, 513))
        (0.1 * loss).backward()
        lay = _lib.state_layout(2, 19)
        accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda")
        out = torch.empty(lay.out_bytes, dtype=torch.uint8, device="cuda")
        aux = torch.empty(lib.msq_fused_aux_bytes(2, 257, 513), dtype=torch.uint8, device="cuda")
        grad = torch.full_like(lo, float("nan"))
        st = torch.cuda.current_stream().cuda_stream
        for a in (aux.data_ptr(), None):              # with and without the statistics cache
            _lib.check(lib.msq_fused_fwd_bwd(mode, lo.data_ptr(), 2, 19, 33, 65, 257, 513, 0.2, 0, accum.data_ptr(), out.data_ptr(),
                                             a, None, 0.1, grad.data_ptr(), None, 0, st))
            got = out[lay.loss_off:lay.loss_off + 4].view(torch.float32).item()
            assert got == loss.item()
            _grad_close(grad, x.grad, rtol=1e-5)


@pytest.mark.parametrize("n,C,hw,HW,rows", [(1, 13, (9, 17), (33, 65), 0), (3, 16, (33, 65), (257, 513), 0),
                                             (5, 19, (17, 33), (129, 257), 40), (2, 7, (12, 20), (90, 150), 0),
                                             (2, 19, (65, 129), (512, 1024), 0)])
def test_one_call_step_is_the_two_calls_bit_for_bit(msq, n, C, hw, HW, rows):
    """msq_fused_fwd_bwd runs TWO kernels -- the backward derives the image-wise weights from the forward's class histogram
    itself and carries the finalisation in an extra CTA -- where msq_fused_fwd + msq_fused_bwd run three (forward,
    finalisation, backward).  Same integers, same arithmetic: every output (loss, weights, histogram, per-image sums,
    statistics vector) must be bit-identical, the gradient equal to within the order of its atomic adds, and the
    accumulator buffer all-zero again (self-clean, hand-over counter included) so that it can be reused at once."""
    from maxsquareloss_b200 import _lib
    lib = _lib.load()
    (h, w), (H, W) = hw, HW
    st = torch.cuda.current_stream().cuda_stream
    lay = _lib.state_layout(n, C)
    aux = torch.empty(lib.msq_fused_aux_bytes(n, H, W), dtype=torch.uint8, device="cuda")
    go = torch.full((), 0.37, device="cuda")
    try:
        _lib.tune("fused_rows", rows)                  # rows > 0: few fat CTAs that walk several segments / images each
        for mode in (_lib.MODE_IW, _lib.MODE_MAXSQUARE):
            accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda")
            for rep in range(3):                       # the same accumulator buffer, three steps in a row
                lo = synth.head_logits(n, C, hw, 100 + rep, 4.0).cuda()
                o2, g2 = torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda"), torch.full_like(lo, float("nan"))
                _lib.check(lib.msq_fused_fwd(mode, lo.data_ptr(), n, C, h, w, H, W, None, 0.2, 0, accum.data_ptr(), o2.data_ptr(),
                                             aux.data_ptr(), g2.data_ptr(), st))
                _lib.check(lib.msq_fused_bwd(mode, lo.data_ptr(), n, C, h, w, H, W, 0, o2.data_ptr(), aux.data_ptr(), go.data_ptr(),
                                             g2.data_ptr(), 1, st))
                torch.cuda.synchronize()
                assert not accum.any()
                for late in (1, 0):
                    _lib.tune("late_finalize", late)
                    o1, g1 = torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda"), torch.full_like(lo, float("nan"))
                    _lib.check(lib.msq_fused_fwd_bwd(mode, lo.data_ptr(), n, C, h, w, H, W, 0.2, 0, accum.data_ptr(), o1.data_ptr(),
                                                     aux.data_ptr(), go.data_ptr(), 0.0, g1.data_ptr(), None, 0, st))
                    torch.cuda.synchronize()
                    assert torch.equal(o1, o2), (mode, rep, late)
                    assert not accum.any(), (mode, rep, late)
                    _grad_close(g1, g2, rtol=1e-5)
    finally:
        _lib.tune("fused_rows", 0)
        _lib.tune("late_finalize", 1)


def test_one_call_step_fuzz_against_the_two_calls(msq):
    """Random geometries (ragged widths, 1-row and 1-column heads, 2..32 classes, up to 9 images, grids of a single CTA up to a
    full wave, forced multi-segment CTAs): the two-kernel one-call step must reproduce the three-kernel two-call path bit for
    bit in every output and leave the accumulators zero."""
    from maxsquareloss_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    rng = np.random.RandomState(20260219)
    go = torch.full((), -0.8, device="cuda")
    try:
        for case in range(36):
            n = int(rng.choice([1, 1, 2, 3, 5, 9]))
            C = int(rng.choice([2, 3, 7, 13, 16, 19, 24, 32]))
            h, w = int(rng.randint(1, 24)), int(rng.randint(1, 40))
            H, W = h + int(rng.randint(0, 8 * h + 3)), w + int(rng.randint(0, 8 * w + 3))
            rows = int(rng.choice([0, 0, 0, 1, 7, 33, 1000]))
            mode = _lib.MODE_IW if case % 3 else _lib.MODE_MAXSQUARE
            _lib.tune("fused_rows", rows)
            lay = _lib.state_layout(n, C)
            accum = torch.zeros(lay.accum_bytes, dtype=torch.uint8, device="cuda")
            aux = torch.empty(lib.msq_fused_aux_bytes(n, H, W), dtype=torch.uint8, device="cuda")
            lo = (torch.randn(n, C, h, w, generator=torch.Generator().manual_seed(case)) * 3.0).cuda()
            o2, g2 = torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda"), torch.full_like(lo, float("nan"))
            _lib.check(lib.msq_fused_fwd(mode, lo.data_ptr(), n, C, h, w, H, W, None, 0.2, 0, accum.data_ptr(), o2.data_ptr(),
                                         aux.data_ptr(), g2.data_ptr(), st))
            _lib.check(lib.msq_fused_bwd(mode, lo.data_ptr(), n, C, h, w, H, W, 0, o2.data_ptr(), aux.data_ptr(), go.data_ptr(),
                                         g2.data_ptr(), 1, st))
            o1, g1 = torch.zeros(lay.out_bytes, dtype=torch.uint8, device="cuda"), torch.full_like(lo, float("nan"))
            for _ in range(2):                          # twice on the same accumulator buffer
                _lib.check(lib.msq_fused_fwd_bwd(mode, lo.data_ptr(), n, C, h, w, H, W, 0.2, 0, accum.data_ptr(), o1.data_ptr(),
                                                 aux.data_ptr(), go.data_ptr(), 0.0, g1.data_ptr(), None, 0, st))
            torch.cuda.synchronize()
            what = (case, n, C, (h, w), (H, W), rows, mode)
            assert torch.equal(o1, o2), what
            assert not accum.any(), what
            assert torch.isfinite(g1).all(), what
            scale = g2.abs().max().item()
            assert (g1 - g2).abs().max().item() <= 1e-5 * scale + 1e-12, what
    finally:
        _lib.tune("fused_rows", 0)
