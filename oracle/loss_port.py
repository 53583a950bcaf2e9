"""CPU PyTorch port of the reference's adaptation losses and their call-site
prologue.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows, op for op, what the reference executes so that timing it is timing
the reference's algorithm (it is the ``--impl reference`` / ``cpu_baseline``
arm of ``bench.py``):

* ``maxsquare``           <- ``utils/loss.py:110-119``  (MaxSquareloss.forward)
* ``iw_maxsquare``        <- ``utils/loss.py:76-102``   (IW_MaxSquareloss.forward)
* ``prologue``            <- ``graphs/models/deeplab_multi.py:124,128`` +
                             ``tools/solve_gta5.py:182-183``
* ``chain_*``             <- prologue + loss + ``.backward()``
                             (``tools/solve_gta5.py:199,217``)

One deliberate difference: the reference's IW loss multiplies a (N,C,H,W)
tensor by (N,H,W) weights without ``unsqueeze(1)`` (``utils/loss.py:98-100``)
and therefore raises for N >= 2 (unless N == C).  The port inserts the
``unsqueeze(1)``, which for N == 1 is the identical computation and for N > 1
is the evident intent (per-image weights; mean over images of the N=1 loss).
"""
import torch
import torch.nn.functional as F


def maxsquare(prob: torch.Tensor, ignore_index: int = -1) -> torch.Tensor:
    # utils/loss.py:117-118
    keep = prob != ignore_index
    return -torch.mean(torch.pow(prob, 2)[keep]) / 2


def image_weights_from_hist(hist: torch.Tensor, ratio: float) -> torch.Tensor:
    # utils/loss.py:95  (fp32, CPU): 1 / max(hist^r * total^(1-r), 1)
    return 1 / torch.max(torch.pow(hist, ratio) * torch.pow(hist.sum(), 1 - ratio), torch.ones(1))


def class_hist(label_img: torch.Tensor, num_class: int) -> torch.Tensor:
    # utils/loss.py:92-94: histc over [-1, C-1] with C+1 bins, bin 0 (= -1) dropped
    hist = torch.histc(label_img.cpu().data.float(), bins=num_class + 1, min=-1,
                       max=num_class - 1).float()
    return hist[1:]


def iw_maxsquare(prob: torch.Tensor, num_class: int, ratio: float = 0.2,
                 label=None, ignore_index: int = -1, return_aux: bool = False):
    # utils/loss.py:84-100
    maxpred, argpred = torch.max(prob, 1)
    keep_px = maxpred != ignore_index
    argpred = torch.where(keep_px, argpred,
                          torch.ones(1).to(prob.device, dtype=torch.long) * ignore_index)
    if label is None:
        label = argpred
    n = prob.size(0)
    per_px, hists = [], []
    for i in range(n):
        hist = class_hist(label[i], num_class)
        w = image_weights_from_hist(hist, ratio).to(argpred.device)[argpred[i]].detach()
        per_px.append(w)
        hists.append(hist)
    weights = torch.stack(per_px, dim=0).unsqueeze(1)      # see module docstring
    keep = keep_px.unsqueeze(1).expand_as(prob)
    loss = -torch.sum((torch.pow(prob, 2) * weights)[keep]) / (n * num_class)
    if return_aux:
        return loss, torch.stack(hists).to(torch.int64), argpred
    return loss


def prologue(logits_lo: torch.Tensor, out_hw):
    """low-res head logits -> (pred, prob) exactly as the reference's model and
    trainer produce them."""
    pred = F.interpolate(logits_lo, size=tuple(out_hw), mode='bilinear', align_corners=True)
    prob = F.softmax(pred, dim=1)
    return pred, prob


def chain_iw_maxsquare(logits_lo: torch.Tensor, out_hw, num_class: int, ratio: float = 0.2,
                       grad_scale: float = 1.0):
    """prologue -> IW loss -> backward.  Returns (loss, grad wrt low-res logits, hist)."""
    x = logits_lo.detach().clone().requires_grad_(True)
    _, prob = prologue(x, out_hw)
    loss, hist, _ = iw_maxsquare(prob, num_class, ratio, return_aux=True)
    (grad_scale * loss).backward()
    return loss.detach(), x.grad, hist


def chain_maxsquare(logits_lo: torch.Tensor, out_hw, grad_scale: float = 1.0):
    x = logits_lo.detach().clone().requires_grad_(True)
    _, prob = prologue(x, out_hw)
    loss = maxsquare(prob)
    (grad_scale * loss).backward()
    return loss.detach(), x.grad


def multi_level_guidance(pred: torch.Tensor, pred_2: torch.Tensor, threshold: float,
                         ignore_index: int = -1):
    """Self-produced guidance of the second head (``tools/solve_gta5.py:183,192,
    206-215`` == ``tools/solve_crosscity.py:235-243``): returns (label_2, CE loss
    of pred_2 against it, mean over valid pixels, before the lambda factors)."""
    p1 = F.softmax(pred, dim=1)
    p2 = F.softmax(pred_2, dim=1)
    max1, _ = torch.max(p1.detach(), dim=1)
    max2, _ = torch.max(p2.detach(), dim=1)
    pc = (p1 + p2) / 2
    _, arg_c = torch.max(pc, dim=1)
    keep = (max1 > threshold) | (max2 > threshold)
    label_2 = torch.where(keep, arg_c, torch.ones(1).to(pred.device, dtype=torch.long) * ignore_index)
    loss_2 = F.cross_entropy(pred_2, label_2, ignore_index=ignore_index)
    return label_2, loss_2


def chain_multi(lo1: torch.Tensor, lo2: torch.Tensor, out_hw, num_class: int, kind: str = "iw",
                ratio: float = 0.2, threshold: float = 0.95, lambda_target: float = 0.1,
                lambda_seg: float = 0.1):
    """The whole ``--multi`` target step (``tools/solve_gta5.py:178-218``) from the two heads'
    low-resolution logits: both upsamples, both softmaxes, head-1 adaptation loss, guidance CE on
    head 2, backward.  Returns dict(loss_target, loss_target_2, label_2, nvalid, grad1, grad2, hist)."""
    x1 = lo1.detach().clone().requires_grad_(True)
    x2 = lo2.detach().clone().requires_grad_(True)
    pred, prob = prologue(x1, out_hw)
    pred_2 = F.interpolate(x2, size=tuple(out_hw), mode='bilinear', align_corners=True)
    hist = None
    if kind == "iw":
        loss1, hist, _ = iw_maxsquare(prob, num_class, ratio, return_aux=True)
    else:
        loss1 = maxsquare(prob)
    loss_target = lambda_target * loss1
    label_2, ce = multi_level_guidance(pred, pred_2, threshold)
    loss_target_2 = lambda_seg * lambda_target * ce
    (loss_target + loss_target_2).backward()
    return dict(loss_target=loss_target.detach(), loss_target_2=loss_target_2.detach(), label_2=label_2,
                nvalid=int((label_2 >= 0).sum()), grad1=x1.grad, grad2=x2.grad, hist=hist)


def chain_hard(lo: torch.Tensor, out_hw, threshold: float = 0.95, lambda_target: float = 0.1):
    """``--target_mode hard`` (``tools/solve_gta5.py:149-150,185-199``): pseudo-labels ``argmax(softmax(pred))`` where the
    maximum probability exceeds ``threshold`` (-1 elsewhere), ``lambda_target * CrossEntropyLoss(ignore_index=-1)``,
    backward.  Returns dict(loss_target, label, nvalid, grad)."""
    x = lo.detach().clone().requires_grad_(True)
    pred, prob = prologue(x, out_hw)
    label = torch.argmax(prob.detach(), dim=1)                                   # :185-186
    maxpred, _ = torch.max(prob.detach(), dim=1)                                 # :192
    label = torch.where(maxpred > threshold, label, torch.ones(1, dtype=torch.long) * -1)      # :195-197
    loss = lambda_target * F.cross_entropy(pred, label, ignore_index=-1)         # :199
    loss.backward()
    return dict(loss_target=loss.detach(), label=label, nvalid=int((label >= 0).sum()), grad=x.grad)


def chain_source(logits_lo: torch.Tensor, target: torch.Tensor, num_class: int, grad_scale: float = 1.0):
    """Source-side step (``tools/train_source.py:254,280-283``): upsample -> CrossEntropyLoss(ignore_index=-1)
    -> backward, and the argmax map that goes into ``Eval.add_batch``.
    Returns dict(loss, grad, argpred (numpy int64), nvalid)."""
    import numpy as np
    x = logits_lo.detach().clone().requires_grad_(True)
    pred = F.interpolate(x, size=tuple(target.shape[-2:]), mode='bilinear', align_corners=True)
    loss = F.cross_entropy(pred, target, ignore_index=-1)
    (grad_scale * loss).backward()
    argpred = np.argmax(pred.data.cpu().numpy(), axis=1)
    return dict(loss=loss.detach(), grad=x.grad, argpred=argpred, nvalid=int((target != -1).sum()))


def soft_cross_entropy(inputs: torch.Tensor, target: torch.Tensor, ignore_index: int = -1) -> torch.Tensor:
    # utils/loss.py:29-35 (softCrossEntropy.forward)
    assert inputs.size() == target.size()
    keep = target != ignore_index
    log_likelihood = F.log_softmax(inputs, dim=1)
    return torch.mean(torch.mul(-log_likelihood, target)[keep])


def iw_soft_cross_entropy(inputs: torch.Tensor, target: torch.Tensor, num_class: int, ratio: float = 0.2,
                          ignore_index: int = -1, return_aux: bool = False):
    # utils/loss.py:52-67 (IWsoftCrossEntropy.forward); the unsqueeze(1) is the N > 1 definition
    # (see the module docstring: the reference broadcasts (N,H,W) against (N,C,H,W) and only works for N == 1)
    assert inputs.size() == target.size()
    keep = target != ignore_index
    _, argpred = torch.max(inputs, 1)
    n = inputs.size(0)
    per_px, hists = [], []
    for i in range(n):
        hist = torch.histc(argpred[i].cpu().data.float(), bins=num_class, min=0, max=num_class - 1).float()
        w = image_weights_from_hist(hist, ratio).to(argpred.device)[argpred[i]].detach()
        per_px.append(w)
        hists.append(hist)
    weights = torch.stack(per_px, dim=0).unsqueeze(1)
    log_likelihood = F.log_softmax(inputs, dim=1)
    loss = torch.sum((torch.mul(-log_likelihood, target) * weights)[keep]) / (n * num_class)
    if return_aux:
        return loss, torch.stack(hists).to(torch.int64)
    return loss


def chain_entropy(logits_lo: torch.Tensor, out_hw, num_class: int, iw: bool, ratio: float = 0.2, grad_scale: float = 1.0):
    """prologue -> MinEnt loss called as the trainers call it (target = softmax(pred), attached) -> backward."""
    x = logits_lo.detach().clone().requires_grad_(True)
    pred, prob = prologue(x, out_hw)
    hist = None
    if iw:
        loss, hist = iw_soft_cross_entropy(pred, prob, num_class, ratio, return_aux=True)
    else:
        loss = soft_cross_entropy(pred, prob)
    (grad_scale * loss).backward()
    return loss.detach(), x.grad, hist
