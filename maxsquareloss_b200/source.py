"""Source-side segmentation loss + evaluation of every training iteration, fused from the
LOW-resolution head logits (the other half of the loop the adaptation losses sit in).

The reference does, per iteration (``tools/train_source.py:254-283`` == ``tools/solve_gta5.py:220-235``,
``tools/solve_crosscity.py:183-197``)::

    pred      = model(x)[0]                                   # upsampled inside the model, deeplab_multi.py:124,128
    cur_loss  = nn.CrossEntropyLoss(ignore_index=-1)(pred, y) # train_source.py:128,254
    ...
    argpred   = np.argmax(pred.data.cpu().numpy(), axis=1)    # 40 MB/image D2H + single-thread argmax
    self.Eval.add_batch(y.cpu().numpy(), argpred)

``CrossEntropyLoss2d`` keeps ``nn.CrossEntropyLoss``'s call ``loss(pred, y)`` but ``pred`` may be the
LOW-resolution head logits: the upsample to ``y``'s size, log-softmax, NLL, the argmax and the
confusion-matrix update all happen in one kernel (``msq_source_ce_fwd``), and the backward returns
dL/dlogits at h x w (``msq_guidance_bwd``).  If ``pred`` already has ``y``'s size the interpolation is the
identity and the result is ``nn.CrossEntropyLoss``'s.  CUDA only; no fallback.
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from . import _lib
from .guidance import _GuidanceOutputs, _is_sharded, global_ce_mean
from .loss import _accum_buffer, _device_index, _grad_out_ptr, _new_out, _prep_label, _raw_stream, _require_cuda_f32


class _SourceCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, out_size, evaluator, group, sink):
        n, c, h, w = logits.shape
        H, W = out_size
        lo = logits.contiguous()
        lib = _lib.load()
        lay = _lib.state_layout(n, c)
        accum, stream = _accum_buffer(lo.device, lay.accum_bytes)
        out = _new_out(lay, lo.device)
        need = ctx.needs_input_grad[0]
        aux = torch.empty(16 * n * H * W, dtype=torch.uint8, device=lo.device) if need else None
        grad = torch.empty_like(lo) if need else None
        cm_ptr = None
        if evaluator is not None:
            cm_ptr = evaluator._dev.data_ptr()
            evaluator._pending = True
        _lib.check(lib.msq_source_ce_fwd(lo.data_ptr(), target.data_ptr(), n, c, h, w, H, W, accum.data_ptr(), out.data_ptr(),
                                         aux.data_ptr() if need else None, grad.data_ptr() if need else None, cm_ptr, stream))
        o = _GuidanceOutputs(out, n, c, lay)
        loss = o.loss2
        if _is_sharded(group):
            loss = global_ce_mean(o, group)              # the mean is over the valid pixels of the WHOLE batch
        sink.append(o)
        ctx.save_for_backward(lo)
        ctx.keep = (out, aux, grad)
        ctx.cfg = (H, W)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        if not ctx.needs_input_grad[0]:
            return (None,) * 6
        (lo,) = ctx.saved_tensors
        out, aux, grad = ctx.keep
        if grad is None:
            raise RuntimeError("CrossEntropyLoss2d: backward called twice; call forward again")
        ctx.keep = (out, aux, None)
        H, W = ctx.cfg
        n, c, h, w = lo.shape
        go = _grad_out_ptr(grad_out, lo.device)
        stream = _raw_stream(_device_index(lo.device))
        _lib.check(_lib.load().msq_guidance_bwd(lo.data_ptr(), n, c, h, w, H, W, out.data_ptr(), aux.data_ptr(), go.data_ptr(),
                                                grad.data_ptr(), 1, stream))
        return (grad,) + (None,) * 5


class CrossEntropyLoss2d(nn.Module):
    """``nn.CrossEntropyLoss(weight=None, ignore_index=-1)`` (``tools/train_source.py:128``) on head logits
    that are bilinearly upsampled (``align_corners=True``) to the target's size inside the kernel.

    :param ignore_index: only -1 (what every trainer of the reference passes) is supported; labels
                         outside ``[0, C)`` are ignored as well
    :param evaluator:    optional ``maxsquareloss_b200.Eval``: every forward also accumulates the
                         confusion matrix of ``argmax(pred)`` against ``target`` into it
                         (``tools/train_source.py:280-283``) at no extra pass over the data
    """

    def __init__(self, weight=None, ignore_index=-1, evaluator=None, group=None):
        super().__init__()
        if weight is not None:
            raise RuntimeError("class weights are not supported (the reference passes weight=None)")
        if ignore_index != -1:
            raise RuntimeError("only ignore_index=-1 is supported (the reference's value)")
        self.ignore_index = ignore_index
        self.evaluator = evaluator
        self.group = group
        self.last_nvalid = None
        self.last_ce_sum = None

    def forward(self, pred, target, evaluator=None):
        _require_cuda_f32(pred, "pred")
        if pred.shape[1] > _lib.MAX_CLASSES:
            raise RuntimeError(f"{pred.shape[1]} classes exceed the kernels' limit of {_lib.MAX_CLASSES}")
        if not isinstance(target, torch.Tensor) or target.dim() != 3 or target.shape[0] != pred.shape[0]:
            raise RuntimeError("target must be an (N,H,W) integer tensor")
        H, W = int(target.shape[1]), int(target.shape[2])
        tgt = _prep_label(target, pred.device, (pred.shape[0], H, W), "target")
        ev = evaluator if evaluator is not None else self.evaluator
        if ev is not None and ev.num_class != pred.shape[1]:
            raise ValueError(f"pred has {pred.shape[1]} classes, the evaluator was built with {ev.num_class}")
        sink = []
        loss = _SourceCE.apply(pred, tgt, (H, W), ev, self.group, sink)
        self.last_nvalid, self.last_ce_sum = sink[0].nvalid, sink[0].ce_sum
        return loss
