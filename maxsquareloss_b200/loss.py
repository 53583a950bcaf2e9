"""Host side of the adaptation losses: the reference's ``nn.Module`` API over the
sm_100a kernels in ``csrc/`` (C ABI: ``include/msq_b200.h``).

Drop-in for ``utils/loss.py`` of shiyutang/MaxSquareLoss:

* ``MaxSquareloss(ignore_index=-1, num_class=19).forward(pred, prob)``
  (``utils/loss.py:104-119``)
* ``IW_MaxSquareloss(ignore_index=-1, num_class=19, ratio=0.2).forward(pred, prob, label=None)``
  (``utils/loss.py:69-102``)

Two entry modes:

strict   ``forward(pred, prob[, label])`` exactly as the trainers call it
         (``tools/solve_gta5.py:199``): ``prob`` is the full-resolution softmax
         output, ``pred`` is ignored (the reference ignores it too), the gradient
         is returned w.r.t. ``prob``.
fused    ``forward(head_logits, out_size=(H, W)[, label=...])`` with ``prob`` omitted:
         ``head_logits`` are the LOW-resolution classifier outputs (N,C,h,w); the
         model's ``F.interpolate(..., 'bilinear', align_corners=True)``
         (``graphs/models/deeplab_multi.py:124,128``) and the trainer's
         ``F.softmax`` are done inside the kernels and the gradient is returned
         w.r.t. the low-resolution logits.

Differences from the reference, all deliberate:

* the IW loss accepts N >= 2 (the reference raises, ``utils/loss.py:98-100`` lacks
  an ``unsqueeze(1)``): per-image weights, i.e. the mean over images of the N=1 loss;
* nothing synchronises the host (the reference does a D2H + CPU histc + H2D per image);
* CUDA fp32 tensors only -- ``RuntimeError`` otherwise; there is no CPU fallback.
"""
import torch
import torch.nn as nn

from . import _lib, _torch_ops

_accum_cache = {}

#: fused mode: let the forward write the 16 B/pixel statistics cache for the backward
USE_STATS_CACHE = True

try:                                        # raw cudaStream_t of the current stream without building a torch.cuda.Stream
    _raw_stream = torch._C._cuda_getCurrentRawStream          # (torch.cuda.current_stream() costs ~12 us per call)
except AttributeError:                      # pragma: no cover
    def _raw_stream(index):
        return torch.cuda.current_stream(index).cuda_stream


def _device_index(device):
    idx = device.index
    return idx if idx is not None else torch.cuda.current_device()


def _accum_buffer(device, nbytes):
    """Zero-initialised, self-cleaning accumulator buffer, one per (device, stream)."""
    idx = _device_index(device)
    stream = _raw_stream(idx)
    key = (idx, stream)
    buf = _accum_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(int(nbytes), 4096), dtype=torch.uint8, device=device)
        _accum_cache[key] = buf
    return buf, stream


def reset_workspaces():
    """Drop the cached accumulator buffers (only needed after a CUDA error)."""
    _accum_cache.clear()


def _require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError(f"{name} must be a torch.Tensor")
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: maxsquareloss_b200 has no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"{name} must be float32, got {t.dtype}")
    if t.dim() != 4:
        raise RuntimeError(f"{name} must be (N,C,H,W), got shape {tuple(t.shape)}")


def _prep_label(label, device, shape, name):
    if label is None:
        return None
    if not isinstance(label, torch.Tensor):
        label = torch.as_tensor(label)
    label = label.to(device=device, dtype=torch.int64, non_blocking=True)
    if tuple(label.shape) != tuple(shape):
        raise RuntimeError(f"{name} must have shape {tuple(shape)}, got {tuple(label.shape)}")
    return label.contiguous()


class _Outputs:
    """The per-call output buffer of a forward (``msq_state_layout``: every member is 4-byte aligned, so the buffer
    is allocated as float32 words).  Only ``loss`` is materialised eagerly (one indexing op); the other views are
    built when somebody reads them -- a training step reads none."""
    __slots__ = ("buf", "n", "c", "lay", "loss")

    def __init__(self, buf, n, c, lay=None):
        self.buf, self.n, self.c = buf, n, c
        self.lay = lay if lay is not None else _lib.state_layout(n, c)
        self.loss = buf[self.lay.loss_off >> 2]

    def _view(self, off, nbytes, dtype):
        return self.buf[off >> 2:(off + nbytes) >> 2].view(dtype)

    @property
    def weights(self):
        return self._view(self.lay.weights_off, 4 * self.n * self.c, torch.float32).view(self.n, self.c)

    @property
    def hist(self):
        return self._view(self.lay.hist_out_off, 4 * self.n * self.c, torch.int32).view(self.n, self.c)

    @property
    def sum_q(self):
        return self._view(self.lay.sum_out_off, 8 * self.n, torch.float64)

    @property
    def stats(self):
        return self._view(self.lay.stats_off, 8 * (1 + self.c), torch.float64)


def _new_out(lay, device):
    return torch.empty(lay.out_bytes >> 2, dtype=torch.float32, device=device)


def _grad_out_ptr(grad_out, device):
    g = grad_out
    if g.device != device or g.dtype != torch.float32:
        g = g.to(device=device, dtype=torch.float32)
    return g.contiguous()


class _LossBase(nn.Module):
    _mode = None

    def __init__(self, ignore_index=-1, num_class=19):
        super().__init__()
        self.ignore_index = ignore_index
        self.num_class = num_class
        #: normaliser N of utils/loss.py:100 when the batch is sharded by image over
        #: ranks: set to the GLOBAL batch size (0 = this call's own N)
        self.global_batch = 0
        self.__dict__["_last"] = None      # outputs of the most recent forward (plain attribute: nn.Module.__setattr__ costs 5 us)

    # device tensors of the most recent forward (no host sync to produce them; views are built on access)
    def _outputs(self):
        o = self.__dict__["_last"]
        if o is None or isinstance(o, _Outputs):
            return o
        o = _Outputs(*o)
        self.__dict__["_last"] = o
        return o

    @property
    def last_hist(self):
        """(N,C) int32 per-image argmax/label histogram (IW)"""
        o = self._outputs()
        return None if o is None else o.hist

    @property
    def last_weights(self):
        """(N,C) float32 image-wise class weights (IW)"""
        o = self._outputs()
        return None if o is None else o.weights

    @property
    def last_sum_q(self):
        """(N,) float64 per-image sum over pixels of sum_c p_c^2"""
        o = self._outputs()
        return None if o is None else o.sum_q

    @property
    def last_stats(self):
        """(1+C,) float64 [loss, class histogram summed over images]: the vector to all-reduce over ranks"""
        o = self._outputs()
        return None if o is None else o.stats

    def _check_classes(self, c):
        if c != self.num_class:
            raise ValueError(f"tensor has {c} classes but the loss was built with num_class={self.num_class}")
        if c > _lib.MAX_CLASSES:
            raise RuntimeError(f"num_class={c} exceeds the kernels' limit of {_lib.MAX_CLASSES}")

    def _publish(self, sink):
        self.__dict__["_last"] = sink[0]

    def _run(self, pred, prob, label, out_size, ratio, kind=0):
        """One forward through the C++ autograd node (``torch.ops.msq_b200.*``, csrc/torch_binding.cpp): the checks that
        need no Python (device, dtype, rank, label shape) are made there."""
        ops = _torch_ops.load()
        if prob is None:
            if out_size is None:
                raise RuntimeError("fused mode needs out_size=(H, W): forward(head_logits, out_size=...)")
            if not isinstance(pred, torch.Tensor):
                raise RuntimeError("head logits must be a torch.Tensor")
            c = pred.shape[1] if pred.dim() == 4 else -1
            if c != self.num_class and c >= 0 and self.num_class is not None:
                self._check_classes(c)
            # a softmax output never equals an ignore value outside [0,1], so the reference's masks
            # (utils/loss.py:85,117) are all-true and the fused kernels do not evaluate them
            if 0.0 <= self.ignore_index <= 1.0:
                raise RuntimeError("fused mode cannot honour an ignore_index inside [0,1]; use the strict mode")
            H, W = int(out_size[0]), int(out_size[1])
            if label is not None:
                label = _prep_label(label, pred.device, (pred.shape[0], H, W), "label")
            loss, out = ops.fused_loss(pred, label, H, W, self._mode, float(ratio), int(self.global_batch), kind, USE_STATS_CACHE)
            n = pred.shape[0]
        else:
            if not isinstance(prob, torch.Tensor):
                raise RuntimeError("prob must be a torch.Tensor")
            c = prob.shape[1] if prob.dim() == 4 else -1
            if c != self.num_class and c >= 0:
                self._check_classes(c)
            if label is not None:
                label = _prep_label(label, prob.device, (prob.shape[0],) + tuple(prob.shape[2:]), "label")
            loss, out = ops.prob_loss(prob, label, self._mode, float(ratio), int(self.ignore_index), int(self.global_batch))
            n = prob.shape[0]
        self.__dict__["_last"] = (out, n, c)
        return loss


class MaxSquareloss(_LossBase):
    """``-mean(prob**2) / 2`` (``utils/loss.py:104-119``)."""
    _mode = _lib.MODE_MAXSQUARE

    def __init__(self, ignore_index=-1, num_class=19):
        super().__init__(ignore_index, num_class)

    def forward(self, pred, prob=None, out_size=None):
        """
        :param pred: predictions (N, C, H, W) -- unused in strict mode, as in the reference;
                     in fused mode (``prob`` omitted) the low-resolution head logits (N, C, h, w)
        :param prob: probability of pred (N, C, H, W)
        :param out_size: (H, W) label resolution, fused mode only
        :return: maximum squares loss (0-dim CUDA tensor)
        """
        return self._run(pred, prob, None, out_size, 0.0)


class IW_MaxSquareloss(_LossBase):
    """Image-wise weighted maximum squares loss (``utils/loss.py:69-102``)."""
    _mode = _lib.MODE_IW

    def __init__(self, ignore_index=-1, num_class=19, ratio=0.2):
        super().__init__(ignore_index, num_class)
        self.ratio = ratio

    def forward(self, pred, prob=None, label=None, out_size=None):
        """
        :param pred: predictions (N, C, H, W) -- unused in strict mode; head logits in fused mode
        :param prob: probability of pred (N, C, H, W)
        :param label(optional): the map for counting label numbers (N, H, W)
        :param out_size: (H, W) label resolution, fused mode only
        :return: maximum squares loss with image-wise weighting factor (0-dim CUDA tensor)
        """
        return self._run(pred, prob, label, out_size, self.ratio)


class _SoftCE(torch.autograd.Function):
    """Strict MinEnt: full-resolution ``inputs`` and an arbitrary ``target`` (``msq_softce_fwd`` / ``msq_softce_bwd``),
    gradients for both (the trainers' target is ``softmax(inputs)``, attached to the graph)."""

    @staticmethod
    def forward(ctx, inputs, target, mode, ratio, ignore_index, n_norm, sink):
        n, c, h, w = inputs.shape
        z, t = inputs.contiguous(), target.contiguous()
        lay = _lib.state_layout(n, c)
        accum, stream = _accum_buffer(z.device, lay.accum_bytes)
        out = _new_out(lay, z.device)
        _lib.check(_lib.load().msq_softce_fwd(mode, z.data_ptr(), t.data_ptr(), n, c, h * w, float(ratio), int(ignore_index),
                                              int(n_norm), accum.data_ptr(), out.data_ptr(), stream))
        o = _Outputs(out, n, c, lay)
        sink.append(o)
        ctx.save_for_backward(z, t)
        ctx.out = out
        ctx.cfg = (mode, ignore_index, n_norm)
        return o.loss

    @staticmethod
    def backward(ctx, grad_out):
        need_z, need_t = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if not (need_z or need_t):
            return (None,) * 7
        z, t = ctx.saved_tensors
        mode, ignore_index, n_norm = ctx.cfg
        n, c, h, w = z.shape
        go = _grad_out_ptr(grad_out, z.device)
        gz = torch.empty_like(z)
        gt = torch.empty_like(t) if need_t else None
        _lib.check(_lib.load().msq_softce_bwd(mode, z.data_ptr(), t.data_ptr(), n, c, h * w, int(ignore_index), int(n_norm),
                                              ctx.out.data_ptr(), go.data_ptr(), gz.data_ptr(),
                                              gt.data_ptr() if need_t else None, _raw_stream(_device_index(z.device))))
        return (gz if need_z else None, gt) + (None,) * 5


class _EntropyBase(_LossBase):
    def _run_entropy(self, inputs, target, out_size, ratio):
        _require_cuda_f32(inputs, "inputs")
        c = inputs.shape[1]
        if c > _lib.MAX_CLASSES:
            raise RuntimeError(f"{c} classes exceed the kernels' limit of {_lib.MAX_CLASSES}")
        if self.num_class is not None:
            self._check_classes(c)
        sink = []
        if target is not None:
            # strict call (utils/loss.py:23-35, 46-67): ANY target distribution is honoured, read from memory, and both
            # arguments receive their gradient (the trainers pass target = softmax(inputs) attached to the graph,
            # tools/solve_gta5.py:188-190,199: autograd then adds the two paths up)
            assert inputs.size() == target.size()                      # utils/loss.py:29,52
            if out_size is not None and tuple(out_size) != tuple(inputs.shape[2:]):
                raise RuntimeError("with a target tensor, inputs must already be at the target's resolution")
            _require_cuda_f32(target, "target")
            if target.device != inputs.device:
                raise RuntimeError("inputs and target must be on the same device")
            loss = _SoftCE.apply(inputs, target, self._mode, ratio, self.ignore_index, self.global_batch, sink)
            self._publish(sink)
            return loss
        return self._run(inputs, None, None, out_size, ratio, kind=1)        # MinEnt variant of the fused kernels


class softCrossEntropy(_EntropyBase):
    """``mean((-log_softmax(inputs) * target)[target != ignore_index])`` (MinEnt when ``target = softmax(inputs)``,
    ``utils/loss.py:17-35``; selected by ``--target_mode entropy``, ``tools/solve_gta5.py:150-151``)."""
    _mode = _lib.MODE_MAXSQUARE

    def __init__(self, ignore_index=-1):
        super().__init__(ignore_index, None)

    def forward(self, inputs, target=None, out_size=None):
        """
        :param inputs: predictions (N, C, H, W); in fused mode the low-resolution head logits (N, C, h, w)
        :param target: target distribution (N, C, H, W): any values (the trainers pass softmax(inputs)); omit it for the
                       fused mode, which computes softmax(up(inputs)) itself
        :param out_size: (H, W) label resolution, fused mode only
        :return: loss
        """
        return self._run_entropy(inputs, target, out_size, 0.0)


class IWsoftCrossEntropy(_EntropyBase):
    """Image-wise weighted MinEnt (``utils/loss.py:37-67``; ``--target_mode IW_entropy``): class weights
    from the per-image histogram of ``argmax(inputs)``, as in ``IW_MaxSquareloss``."""
    _mode = _lib.MODE_IW

    def __init__(self, ignore_index=-1, num_class=19, ratio=0.2):
        super().__init__(ignore_index, num_class)
        self.ratio = ratio

    def forward(self, inputs, target=None, out_size=None):
        """
        :param inputs: predictions (N, C, H, W); in fused mode the low-resolution head logits (N, C, h, w)
        :param target: target distribution (N, C, H, W): any values; omit it for the fused mode
        :param out_size: (H, W) label resolution, fused mode only
        :return: loss with image-wise weighting factor
        """
        return self._run_entropy(inputs, target, out_size, self.ratio)


def maxsquare_from_logits(head_logits, out_size, global_batch=0):
    """Functional fused MaxSquare loss from low-resolution head logits."""
    crit = MaxSquareloss(-1, head_logits.shape[1])
    crit.global_batch = global_batch
    return crit(head_logits, out_size=out_size)


def iw_maxsquare_from_logits(head_logits, out_size, ratio=0.2, label=None, global_batch=0, return_hist=False):
    """Functional fused IW-MaxSquare loss from low-resolution head logits."""
    crit = IW_MaxSquareloss(-1, head_logits.shape[1], ratio)
    crit.global_batch = global_batch
    loss = crit(head_logits, label=label, out_size=out_size)
    return (loss, crit.last_hist) if return_hist else loss
