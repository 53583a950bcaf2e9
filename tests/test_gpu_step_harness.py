"""Config-5-style adaptation step with the fused losses swapped in: a small random-init two-head network stands in for
DeepLabv2-ResNet101 (the backbone is out of scope and stays on cuDNN; what matters here is that the gradients the fused
kernels hand back at h x w drive the SAME parameter update as the reference's full-resolution chain).

  reference step (oracle port on the same GPU):  model -> F.interpolate x2 -> CE(source) ; softmax -> IW loss (+ multi) -> backward
  fused step:                                    model -> low-res heads -> CrossEntropyLoss2d / MultiLevelTargetLoss -> backward
"""
import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from maxsquareloss_b200 import synth

pytestmark = pytest.mark.gpu


class TwoHeadNet(nn.Module):
    def __init__(self, C):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(3, 16, 3, stride=2, padding=1), nn.ReLU(),
                                  nn.Conv2d(16, 32, 3, stride=2, padding=1), nn.ReLU(),
                                  nn.Conv2d(32, 32, 3, stride=2, padding=1), nn.ReLU())
        self.head1 = nn.Conv2d(32, C, 3, padding=1)      # "layer6" (pred[0])
        self.head2 = nn.Conv2d(32, C, 1)                 # "layer5" (pred[1])

    def forward(self, x):
        f = self.body(x)
        return self.head1(f) * 4, self.head2(f) * 4      # LOW-resolution logits; the reference model upsamples them


def _param_grads(model):
    return torch.cat([p.grad.reshape(-1) for p in model.parameters()]).double()


def test_adaptation_step_matches_reference_chain():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as msq
    from oracle import eval_port, loss_port
    C, HW = 13, (96, 160)
    torch.manual_seed(12345)                               # the reference's default seed (train_source.py:755)
    model = TwoHeadNet(C).cuda()
    xs = torch.randn(1, 3, *HW, device="cuda")             # source image + labels
    ys = synth.blocky_labels(1, HW, C, 3, grid=(6, 10)).cuda()
    xt = torch.randn(1, 3, *HW, device="cuda")             # target image
    lam_t, lam_s, thr = 0.1, 0.1, 0.2

    # ---------------- reference step: full-resolution tensors, torch ops (tools/solve_crosscity.py:165-249)
    model.zero_grad()
    p1, p2 = model(xs)
    pred, pred_2 = (F.interpolate(p, size=HW, mode="bilinear", align_corners=True) for p in (p1, p2))
    loss_src = F.cross_entropy(pred, ys, ignore_index=-1) + lam_s * F.cross_entropy(pred_2, ys, ignore_index=-1)
    loss_src.backward()
    ref_arg = np.argmax(pred.data.cpu().numpy(), axis=1)
    port = eval_port.EvalPort(C)
    port.add_batch(ys.cpu().numpy(), ref_arg)
    t1, t2 = model(xt)
    tp, tp2 = (F.interpolate(p, size=HW, mode="bilinear", align_corners=True) for p in (t1, t2))
    loss_t = lam_t * loss_port.iw_maxsquare(F.softmax(tp, 1), C, 0.2)
    _, ce2 = loss_port.multi_level_guidance(tp, tp2, thr)
    loss_t2 = lam_s * lam_t * ce2
    (loss_t + loss_t2).backward()
    ref_grads = _param_grads(model)
    ref_vals = (loss_src.item(), loss_t.item(), loss_t2.item())

    # ---------------- fused step: nothing at label resolution but the label map
    model.zero_grad()
    ev = msq.Eval(C)
    ce = msq.CrossEntropyLoss2d(ignore_index=-1)
    p1, p2 = model(xs)
    f_src = ce(p1, ys, evaluator=ev) + lam_s * ce(p2, ys)
    f_src.backward()
    multi = msq.MultiLevelTargetLoss(msq.IW_MaxSquareloss(-1, C, 0.2), threshold=thr, lambda_target=lam_t, lambda_seg=lam_s)
    f_t, f_t2 = multi(model(xt), HW)
    (f_t + f_t2).backward()
    got_grads = _param_grads(model)

    for a, b in zip((f_src.item(), f_t.item(), f_t2.item()), ref_vals):
        assert abs(a - b) <= 1e-5 * abs(b)
    assert (got_grads - ref_grads).abs().max().item() <= 1e-4 * ref_grads.abs().max().item()
    assert (got_grads - ref_grads).norm().item() <= 1e-4 * ref_grads.norm().item()
    assert np.array_equal(ev.confusion_matrix, port.confusion_matrix)
    assert ev.Mean_Intersection_over_Union() == port.Mean_Intersection_over_Union()
    # one optimiser step from either set of gradients lands on the same weights
    opt = torch.optim.SGD(model.parameters(), lr=2.5e-4, momentum=0.9, weight_decay=5e-4)   # train_source.py:139-146
    opt.step()
    assert all(torch.isfinite(p).all() for p in model.parameters())


def test_crosscity_step_with_deeplabv2_resnet101():
    """BASELINE config 5 as written: random-init DeepLabv2-ResNet101 (harness/deeplabv2.py, the reference's topology;
    cuDNN, out of scope), batch 1, 13 classes, 512x1024, seed 12345 (train_source.py:755); the step of
    tools/solve_crosscity.py:165-249 without --multi: source CE + Eval.add_batch, target IW-MaxSquare (lambda 0.1).
    Reference chain = the model's two F.interpolate calls + torch ops (oracle port of the losses); fused = the
    low-resolution heads straight into CrossEntropyLoss2d / IW_MaxSquareloss."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import maxsquareloss_b200 as msq
    from harness.deeplabv2 import DeepLabV2Harness, head_size
    from oracle import eval_port, loss_port
    C, HW = 13, (512, 1024)
    tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False            # the reference is plain fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        torch.manual_seed(12345)
        model = DeepLabV2Harness(C).cuda().train()
        xs = torch.randn(1, 3, *HW, device="cuda")
        ys = synth.blocky_labels(1, HW, C, 5).cuda()
        xt = torch.randn(1, 3, *HW, device="cuda")
        lam_t = 0.1

        model.zero_grad()
        pred, _ = model(xs, upsample=True)
        loss_src = F.cross_entropy(pred, ys, ignore_index=-1)
        loss_src.backward()
        port = eval_port.EvalPort(C)
        port.add_batch(ys.cpu().numpy(), np.argmax(pred.data.cpu().numpy(), axis=1))
        tp, _ = model(xt, upsample=True)
        loss_t = lam_t * loss_port.iw_maxsquare(F.softmax(tp, 1), C, 0.2)
        loss_t.backward()
        ref_grads = _param_grads_req(model)
        ref_vals = (loss_src.item(), loss_t.item())

        model.zero_grad()
        ev = msq.Eval(C)
        lo, _ = model(xs)
        assert tuple(lo.shape[2:]) == head_size(*HW) == (65, 129)
        f_src = msq.CrossEntropyLoss2d(ignore_index=-1)(lo, ys, evaluator=ev)
        f_src.backward()
        lt, _ = model(xt)
        f_t = lam_t * msq.IW_MaxSquareloss(-1, C, 0.2)(lt, out_size=HW)
        f_t.backward()
        got_grads = _param_grads_req(model)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32

    for a, b in zip((f_src.item(), f_t.item()), ref_vals):
        assert abs(a - b) <= 1e-5 * abs(b)
    assert (got_grads - ref_grads).abs().max().item() <= 1e-4 * ref_grads.abs().max().item()
    assert (got_grads - ref_grads).norm().item() <= 1e-4 * ref_grads.norm().item()
    assert np.array_equal(ev.confusion_matrix, port.confusion_matrix)
    assert ev.Mean_Intersection_over_Union() == port.Mean_Intersection_over_Union()


def _param_grads_req(model):
    return torch.cat([p.grad.reshape(-1) for p in model.parameters() if p.requires_grad and p.grad is not None]).double()
