"""Multi-level self-produced guidance ("MaxSquare+IW+Multi", BASELINE config 3).

The reference has no class for it: it is inline trainer code, ``UDATrainer.train_target``
(``tools/solve_gta5.py:178-218``) == ``tools/solve_crosscity.py:209-245``, executed when
``--multi`` is given::

    pred_P, pred_P_2 = softmax(pred), softmax(pred_2)                     # both heads, full resolution
    loss_target   = lambda_target * target_loss(pred, pred_P)             # MaxSquare / IW-MaxSquare
    pred_c        = (pred_P + pred_P_2) / 2
    label_2       = where(max(pred_P) > thr | max(pred_P_2) > thr, argmax(pred_c), -1)
    loss_target_2 = lambda_seg * lambda_target * CrossEntropyLoss(ignore_index=-1)(pred_2, label_2)
    (loss_target + loss_target_2).backward()

``MultiLevelTargetLoss`` is that method as an ``nn.Module`` over ONE fused forward kernel
(``msq_multi_fwd``) fed with the LOW-resolution outputs of both classifier heads (the model's two
``F.interpolate`` calls, ``graphs/models/deeplab_multi.py:124,128``, are absorbed) and two backward
kernels (``msq_fused_bwd`` for head 1, ``msq_guidance_bwd`` for head 2).  CUDA only; no fallback.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .loss import (_accum_buffer, _device_index, _grad_out_ptr, _LossBase, _new_out, _Outputs, _raw_stream, _require_cuda_f32)


class _GuidanceOutputs(_Outputs):
    __slots__ = ("loss2",)

    def __init__(self, buf, n, c, lay=None):
        super().__init__(buf, n, c, lay)
        self.loss2 = buf[self.lay.loss2_off >> 2]

    @property
    def nvalid(self):
        return self._view(self.lay.nvalid_out_off, 8, torch.int64).reshape(())

    @property
    def ce_sum(self):
        return self._view(self.lay.ce_out_off, 8, torch.float64).reshape(())

    @property
    def ce_pair(self):
        """(2,) int64 view of the adjacent ``[ce_fix_out | nvalid_out]`` words: the cross-entropy sum as the 2^-32
        fixed-point integer it was accumulated in, and the valid-pixel count"""
        return self._view(self.lay.ce_fix_out_off, 16, torch.int64)


def _is_sharded(group):
    if group is False:
        return False
    if hasattr(group, "allreduce_u64"):                       # dist.StatsComm
        return group.world > 1
    return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1


def global_ce_mean(o, group):
    """Images sharded over ranks: ``CrossEntropyLoss`` averages over the valid pixels of the WHOLE batch, so the
    ``[ce_fix | nvalid]`` pair is all-reduced IN PLACE -- integers: exact, the same bits whatever the sharding of the
    rows' fixed-point partial sums -- and the backward (``msq_guidance_bwd``) then divides by the global count.
    ``group``: a ``torch.distributed`` group (ProcessGroupNCCL) or a ``dist.StatsComm`` (the library's own communicator:
    one ``ncclAllReduce`` on a side stream, ~3 us of host time).  Returns the global mean as a 0-dim float32 tensor."""
    pair = o.ce_pair
    if hasattr(group, "sum_u64_begin"):                      # dist.StatsComm: NVLink mailboxes (or the library's ncclAllReduce)
        group.sum_u64_begin(pair)
        group.sum_u64_end(pair)
    else:
        dist.all_reduce(pair, group=group)
    ce = pair[0].to(torch.float64) * (2.0 ** -32)
    o.ce_sum.copy_(ce)
    return (ce / pair[1]).to(torch.float32)


class _MultiLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits1, logits2, out_size, mode, ratio, threshold, n_norm, want_label, group, sink):
        n, c, h, w = logits1.shape
        H, W = int(out_size[0]), int(out_size[1])
        lo1, lo2 = logits1.contiguous(), logits2.contiguous()
        lib = _lib.load()
        lay = _lib.state_layout(n, c)
        accum, stream = _accum_buffer(lo1.device, lay.accum_bytes)
        out = _new_out(lay, lo1.device)
        need1, need2 = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        nbytes = 16 * n * H * W                     # msq_fused_aux_bytes
        aux1 = torch.empty(nbytes, dtype=torch.uint8, device=lo1.device) if need1 else None
        aux2 = torch.empty(nbytes, dtype=torch.uint8, device=lo1.device) if need2 else None
        g1 = torch.empty_like(lo1) if need1 else None
        g2 = torch.empty_like(lo2) if need2 else None
        label2 = torch.empty((n, H, W), dtype=torch.int64, device=lo1.device) if want_label else None
        ptr = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        _lib.check(lib.msq_multi_fwd(mode, lo1.data_ptr(), lo2.data_ptr(), n, c, h, w, H, W, float(ratio),
                                     float(threshold), int(n_norm), accum.data_ptr(), out.data_ptr(), ptr(aux1), ptr(aux2),
                                     ptr(g1), ptr(g2), ptr(label2), stream))
        o = _GuidanceOutputs(out, n, c, lay)
        loss2 = o.loss2
        if _is_sharded(group):
            loss2 = global_ce_mean(o, group)
        sink.append((o, label2))
        ctx.save_for_backward(lo1, lo2)
        ctx.keep = (out, aux1, aux2, g1, g2)
        ctx.cfg = (mode, H, W, n_norm)
        ctx.mark_non_differentiable(*([label2] if label2 is not None else []))
        if label2 is not None:
            return o.loss, loss2, label2
        return o.loss, loss2

    @staticmethod
    def backward(ctx, go1, go2, *unused):
        lo1, lo2 = ctx.saved_tensors
        out, aux1, aux2, g1, g2 = ctx.keep
        ctx.keep = (out, aux1, aux2, None, None)       # the pre-zeroed buffers are good for one backward
        mode, H, W, n_norm = ctx.cfg
        n, c, h, w = lo1.shape
        lib = _lib.load()
        stream = _raw_stream(_device_index(lo1.device))
        r1 = r2 = None
        if ctx.needs_input_grad[0]:
            if g1 is None:
                raise RuntimeError("MultiLevelTargetLoss: backward through head 1 twice")
            go = _grad_out_ptr(go1, lo1.device)
            _lib.check(lib.msq_fused_bwd(mode, lo1.data_ptr(), n, c, h, w, H, W, int(n_norm), out.data_ptr(),
                                         aux1.data_ptr(), go.data_ptr(), g1.data_ptr(), 1, stream))
            r1 = g1
        if ctx.needs_input_grad[1]:
            if g2 is None:
                raise RuntimeError("MultiLevelTargetLoss: backward through head 2 twice")
            go = _grad_out_ptr(go2, lo2.device)
            _lib.check(lib.msq_guidance_bwd(lo2.data_ptr(), n, c, h, w, H, W, out.data_ptr(), aux2.data_ptr(),
                                            go.data_ptr(), g2.data_ptr(), 1, stream))
            r2 = g2
        return (r1, r2) + (None,) * 8


class MultiLevelTargetLoss(nn.Module):
    """``UDATrainer.train_target`` with ``--multi`` (``tools/solve_gta5.py:178-218``), fused.

    :param target_loss: a ``MaxSquareloss`` or ``IW_MaxSquareloss`` of this package (it supplies
                        the mode, ``num_class``, ``ratio`` and ``global_batch``), as the trainer's
                        ``self.target_loss`` (``tools/solve_gta5.py:156-160``)
    :param threshold:   ``--threshold`` (0.95, ``tools/solve_gta5.py:428``; 0.98 crosscity)
    :param lambda_target, lambda_seg: ``--lambda_target`` / ``--lambda_seg`` (0.1, ``tools/train_source.py:827``)

    ``forward((head1, head2), out_size)`` takes the two heads' LOW-resolution logits in the order
    the model returns them (``deeplab_multi.py:130``: ``pred[0]`` is the layer-4 head) and returns
    ``(loss_target, loss_target_2)`` already scaled like the trainer's
    ``self.loss_target`` / ``self.loss_target_2``; ``(loss_target + loss_target_2).backward()``
    then yields both heads' dL/dlogits at h x w.  ``last_label_2`` holds the pseudo-label map when
    ``return_label=True``; ``last_nvalid`` the number of pixels with ``label_2 != -1``.
    """

    def __init__(self, target_loss, threshold=0.95, lambda_target=0.1, lambda_seg=0.1, return_label=False,
                 group=None):
        super().__init__()
        if not isinstance(target_loss, _LossBase):
            raise TypeError("target_loss must be a maxsquareloss_b200 MaxSquareloss / IW_MaxSquareloss")
        self.target_loss = target_loss
        self.threshold = threshold
        self.lambda_target = lambda_target
        self.lambda_seg = lambda_seg
        self.return_label = return_label
        self.group = group            # process group or dist.StatsComm for the sharded CE mean (False = never all-reduce)
        self.ignore_index = -1
        self.last_label_2 = None
        self.last_nvalid = None
        self.last_ce_sum = None
        self.loss_target = None
        self.loss_target_2 = None

    def forward(self, pred, out_size):
        if not isinstance(pred, (tuple, list)) or len(pred) != 2:
            raise RuntimeError("MultiLevelTargetLoss needs the two heads' logits: forward((pred, pred_2), out_size)")
        head1, head2 = pred
        _require_cuda_f32(head1, "head 1 logits")
        _require_cuda_f32(head2, "head 2 logits")
        if head1.shape != head2.shape or head1.device != head2.device:
            raise RuntimeError(f"the two heads must have the same shape and device, got {tuple(head1.shape)} and "
                               f"{tuple(head2.shape)}")
        tl = self.target_loss
        tl._check_classes(head1.shape[1])
        sink = []
        res = _MultiLoss.apply(head1, head2, tuple(out_size), tl._mode, getattr(tl, "ratio", 0.0), self.threshold,
                               tl.global_batch, self.return_label, self.group, sink)
        o, label2 = sink[0]
        tl._publish([o])
        self.last_label_2, self.last_nvalid, self.last_ce_sum = label2, o.nvalid, o.ce_sum
        self.loss_target = self.lambda_target * res[0]
        self.loss_target_2 = self.lambda_seg * self.lambda_target * res[1]
        return self.loss_target, self.loss_target_2


class HardTargetLoss(nn.Module):
    """``--target_mode hard`` of ``UDATrainer.train_target`` (``tools/solve_gta5.py:149-150,185-199``), fused:
    ``label = where(max softmax(pred) > threshold, argmax softmax(pred), -1)`` and
    ``lambda_target * nn.CrossEntropyLoss(ignore_index=-1)(pred, label)`` from the LOW-resolution head logits.

    It is the guidance kernel with both heads bound to the same tensor: ``(P + P) / 2 == P`` exactly, so the ensemble
    argmax / threshold test of ``msq_multi_fwd`` is this mode's pseudo-label and its head-2 cross-entropy is this loss;
    ``msq_guidance_bwd`` returns the gradient.  (One softmax more than a dedicated kernel would need; this mode is the
    reference's self-training baseline, not the path the paper's numbers use.)  CUDA only; no fallback.

    ``forward(pred, out_size)`` returns ``self.loss_target`` (already scaled by ``lambda_target``);
    ``last_label`` holds the pseudo-label map when ``return_label=True``, ``last_nvalid`` the number of kept pixels."""

    def __init__(self, threshold=0.95, lambda_target=0.1, num_class=19, return_label=False, group=None):
        super().__init__()
        self.threshold, self.lambda_target, self.num_class = threshold, lambda_target, num_class
        self.return_label, self.group = return_label, group
        self.ignore_index = -1
        self.last_label = self.last_nvalid = self.loss_target = None

    def forward(self, pred, out_size):
        if isinstance(pred, (tuple, list)):
            pred = pred[0]
        _require_cuda_f32(pred, "head logits")
        if pred.shape[1] != self.num_class:
            raise ValueError(f"tensor has {pred.shape[1]} classes but the loss was built with num_class={self.num_class}")
        if pred.shape[1] > _lib.MAX_CLASSES:
            raise RuntimeError(f"num_class={pred.shape[1]} exceeds the kernels' limit of {_lib.MAX_CLASSES}")
        sink = []
        # head 1 of the kernel is a detached alias (its maximum-squares statistics are not used and get no gradient)
        res = _MultiLoss.apply(pred.detach(), pred, tuple(out_size), _lib.MODE_MAXSQUARE, 0.0, self.threshold, 0,
                               self.return_label, self.group, sink)
        o, label = sink[0]
        self.last_label, self.last_nvalid = label, o.nvalid
        self.loss_target = self.lambda_target * res[1]
        return self.loss_target
