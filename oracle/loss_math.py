"""Closed-form float64 restatement of the two losses and their gradients,
independent of autograd.  TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Derivation (checked against the port + autograd in
``tests/test_oracle_golden.py``):

MaxSquare (``utils/loss.py:117-118``), all elements kept (prob never == -1):
    L        = - sum(p^2) / (2 * M),            M = N*C*H*W
    dL/dp    = - p / M
    dL/dz_c  = - (1/M) * p_c * (p_c - q),       q = sum_k p_k^2   (through softmax)

IW-MaxSquare (``utils/loss.py:84-100``), weights detached:
    k(px)    = argmax_c p_c        (first maximum, torch.max / np.argmax rule)
    hist_n   = bincount of k over image n          (``label`` overrides the counted map)
    w_{n,j}  = 1 / max(hist_{n,j}^r * (sum_j hist_{n,j})^(1-r), 1)      [fp32]
    L        = - (1/(N*C)) * sum_n sum_px w_{n,k(px)} * q(px)
    dL/dp_c  = - 2 * w_{n,k(px)} * p_c / (N*C)
    dL/dz_c  = - (2 * w_{n,k(px)} / (N*C)) * p_c * (p_c - q)

Gradients w.r.t. the low-resolution head logits are the bilinear adjoint of
dL/dz (``oracle/bilinear.py``).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import bilinear
from .loss_port import image_weights_from_hist


def softmax64(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.float64)
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def argmax_of_prob_fp32(z32: np.ndarray) -> np.ndarray:
    """What the reference counts: argmax over the *fp32 softmax output*
    (``utils/loss.py:84``), first maximum wins."""
    p = F.softmax(torch.from_numpy(np.ascontiguousarray(z32)), dim=1)
    return torch.max(p, 1)[1].numpy()


def class_hist_np(k: np.ndarray, num_class: int) -> np.ndarray:
    """(N,H,W) int -> (N,C) int64; values outside [0,C-1] are not counted
    (``utils/loss.py:92-94``: bin 0 of the C+1 histc bins holds -1 and is
    dropped; out-of-range values fall outside [min,max])."""
    out = np.zeros((k.shape[0], num_class), dtype=np.int64)
    for n in range(k.shape[0]):
        v = k[n].reshape(-1)
        v = v[(v >= 0) & (v < num_class)]
        out[n] = np.bincount(v, minlength=num_class)
    return out


def weights_fp32(hist: np.ndarray, ratio: float) -> np.ndarray:
    """(N,C) counts -> (N,C) float32 weights, computed with the reference's own
    fp32 torch expression (``utils/loss.py:95``)."""
    return np.stack([image_weights_from_hist(torch.from_numpy(h.astype(np.float32)), ratio).numpy()
                     for h in hist]).astype(np.float32)


def iw_from_prob(p: np.ndarray, num_class: int, ratio: float, k=None, hist=None):
    """p: (N,C,H,W) float64 probabilities.  Returns dict(loss, grad_prob, hist, w, k, q)."""
    N, C = p.shape[:2]
    if k is None:
        k = p.argmax(axis=1)
    if hist is None:
        hist = class_hist_np(k, num_class)
    w = weights_fp32(hist, ratio).astype(np.float64)
    wpx = np.take_along_axis(w[:, :, None, None], k[:, None, :, :], axis=1)  # (N,1,H,W)
    q = (p * p).sum(axis=1, keepdims=True)
    loss = -(wpx * q).sum() / (N * num_class)
    grad_prob = -2.0 * wpx * p / (N * num_class)
    return dict(loss=loss, grad_prob=grad_prob, hist=hist, w=w, k=k, q=q, wpx=wpx)


def ms_from_prob(p: np.ndarray):
    M = p.size
    q = (p * p).sum(axis=1, keepdims=True)
    return dict(loss=-q.sum() / (2.0 * M), grad_prob=-p / M, q=q)


def fused_iw(lo32: np.ndarray, out_hw, num_class: int, ratio: float = 0.2, grad_scale: float = 1.0):
    """low-res logits -> loss, hist, grad wrt low-res logits (float64)."""
    z = bilinear.upsample(lo32, out_hw)
    p = softmax64(z)
    k = argmax_of_prob_fp32(z)
    r = iw_from_prob(p, num_class, ratio, k=k)
    N = p.shape[0]
    gz = -(2.0 * r['wpx'] / (N * num_class)) * p * (p - r['q']) * grad_scale
    r['grad_logits'] = bilinear.upsample_adjoint(gz, lo32.shape[2:])
    r['z'] = z
    return r


def fused_ms(lo32: np.ndarray, out_hw, grad_scale: float = 1.0):
    z = bilinear.upsample(lo32, out_hw)
    p = softmax64(z)
    r = ms_from_prob(p)
    gz = -(1.0 / p.size) * p * (p - r['q']) * grad_scale
    r['grad_logits'] = bilinear.upsample_adjoint(gz, lo32.shape[2:])
    r['z'] = z
    return r


def near_tie_pixels(z32: np.ndarray, ulps: int = 16) -> np.ndarray:
    """Boolean (N,H,W) map of pixels whose top-2 interpolated logits are within
    ``ulps`` fp32 ulps of each other -- the only pixels where an argmax can
    legitimately differ between arithmetics (hard parts)."""
    s = np.sort(z32, axis=1)
    top, second = s[:, -1], s[:, -2]
    return (top - second) <= ulps * np.spacing(np.abs(top).astype(np.float32))


def guidance(lo1_32: np.ndarray, lo2_32: np.ndarray, out_hw, threshold: float, grad_scale: float = 1.0):
    """Closed form of the multi-level guidance term (``tools/solve_gta5.py:206-213``) in float64:
    label_2 from the fp32 softmaxes exactly as torch computes them, CE and its gradient w.r.t. the
    head-2 low-resolution logits without autograd:
        loss2      = mean over valid px of -log p2[label_2]
        dloss2/dz2 = (p2 - onehot(label_2)) / n_valid  on valid px, 0 elsewhere."""
    z1 = bilinear.upsample(lo1_32, out_hw)
    z2 = bilinear.upsample(lo2_32, out_hw)
    p1 = F.softmax(torch.from_numpy(z1), dim=1)
    p2 = F.softmax(torch.from_numpy(z2), dim=1)
    keep = (torch.max(p1, 1)[0] > threshold) | (torch.max(p2, 1)[0] > threshold)
    lab = torch.where(keep, torch.max((p1 + p2) / 2, 1)[1], torch.full((1,), -1, dtype=torch.long)).numpy()
    valid = lab >= 0
    nvalid = int(valid.sum())
    p64 = softmax64(z2)
    z64 = z2.astype(np.float64)
    lse = np.log(np.exp(z64 - z64.max(axis=1, keepdims=True)).sum(axis=1)) + z64.max(axis=1)
    sel = np.take_along_axis(z64, np.maximum(lab, 0)[:, None], axis=1)[:, 0]
    with np.errstate(all='ignore'):
        loss2 = np.float64((lse - sel)[valid].sum()) / nvalid if nvalid else np.float64('nan')
        onehot = np.zeros_like(p64)
        np.put_along_axis(onehot, np.maximum(lab, 0)[:, None], 1.0, axis=1)
        # torch: with zero valid pixels the loss is NaN (0/0) but nll_loss_backward yields zeros
        gz = (p64 - onehot) * valid[:, None] * (grad_scale / nvalid if nvalid else 0.0)
    return dict(label_2=lab, nvalid=nvalid, loss2=loss2, grad_logits2=bilinear.upsample_adjoint(gz, lo2_32.shape[2:]),
                z1=z1, z2=z2)


def source_ce(lo32: np.ndarray, target: np.ndarray, grad_scale: float = 1.0):
    """Closed form (float64) of ``CrossEntropyLoss(ignore_index=-1)`` on the upsampled logits and of its
    gradient w.r.t. the low-resolution logits; plus ``np.argmax`` of the fp32 upsampled logits."""
    z = bilinear.upsample(lo32, target.shape[-2:])
    z64 = z.astype(np.float64)
    valid = (target >= 0) & (target < lo32.shape[1])
    nvalid = int(valid.sum())
    zmax = z64.max(axis=1)
    lse = np.log(np.exp(z64 - zmax[:, None]).sum(axis=1)) + zmax
    tgt = np.where(valid, target, 0)
    sel = np.take_along_axis(z64, tgt[:, None], axis=1)[:, 0]
    loss = np.float64((lse - sel)[valid].sum()) / nvalid if nvalid else np.float64('nan')
    p = softmax64(z)
    onehot = np.zeros_like(p)
    np.put_along_axis(onehot, tgt[:, None], 1.0, axis=1)
    gz = (p - onehot) * valid[:, None] * (grad_scale / nvalid if nvalid else 0.0)
    return dict(loss=loss, nvalid=nvalid, argpred=z.argmax(axis=1), z=z,
                grad_logits=bilinear.upsample_adjoint(gz, lo32.shape[2:]))


def fused_entropy(lo32: np.ndarray, out_hw, num_class: int, iw: bool, ratio: float = 0.2, grad_scale: float = 1.0):
    """Closed form (float64) of softCrossEntropy / IWsoftCrossEntropy with target = softmax(inputs)
    attached (``utils/loss.py:17-67`` as called at ``tools/solve_gta5.py:188-190,199``):
        H_px   = -sum_c p_c log p_c
        L      = sum_px H_px / (N C H W)                 (softCrossEntropy)
        L      = sum_px w[n, argmax_c z] H_px / (N C)    (IW; argmax of the LOGITS, first maximum)
        dL/dz_j = -coef_px * p_j (log p_j + H_px)        (total derivative through inputs and target)"""
    z = bilinear.upsample(lo32, out_hw)
    z64 = z.astype(np.float64)
    p = softmax64(z)
    logp = z64 - z64.max(axis=1, keepdims=True)
    logp = logp - np.log(np.exp(logp).sum(axis=1, keepdims=True))
    Hpx = -(p * logp).sum(axis=1, keepdims=True)
    N = p.shape[0]
    hist = None
    if iw:
        k = z.argmax(axis=1)
        hist = class_hist_np(k, num_class)
        w = weights_fp32(hist, ratio).astype(np.float64)
        coef = np.take_along_axis(w[:, :, None, None], k[:, None, :, :], axis=1) / (N * num_class)
    else:
        coef = np.full_like(Hpx, 1.0 / p.size)
    loss = (coef * Hpx).sum()
    gz = -coef * p * (logp + Hpx) * grad_scale
    return dict(loss=loss, hist=hist, grad_logits=bilinear.upsample_adjoint(gz, lo32.shape[2:]), z=z)
