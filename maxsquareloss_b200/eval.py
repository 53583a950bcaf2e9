"""Host side of the evaluation accumulator: the reference's ``Eval`` API
(``utils/eval.py:14-124``) over the sm_100a confusion-matrix kernels.

``add_batch`` accepts what the reference's callers pass (numpy ``gt``/``pred``
maps, ``tools/train_source.py:280-283,455-459``) and, as the fast path, CUDA
tensors -- including the raw logits ``(N,C,H,W)``, in which case the callers'
``np.argmax(pred, axis=1)`` is fused into the kernel and the 40 MB/image D2H of
the logits disappears.  Counts are accumulated on the device as uint64 and
folded into the float64 ``confusion_matrix`` attribute when it (or any metric)
is read; the metrics themselves are the reference's NumPy expressions on that
float64 matrix, so they are bit-identical.
"""
import ctypes
import warnings

import numpy as np
import torch

from . import _lib
from .loss import _raw_stream

#: class names the reference prints (datasets/cityscapes_Dataset.py:379-400)
name_classes = ['road', 'sidewalk', 'building', 'wall', 'fence', 'pole', 'trafflight', 'traffsign',
                'vegetation', 'terrain', 'sky', 'person', 'rider', 'car', 'truck', 'bus', 'train',
                'motorcycle', 'bicycle', 'unlabeled']

synthia_set_16 = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 15, 17, 18]     # utils/eval.py:10
synthia_set_13 = [0, 1, 2, 6, 7, 8, 10, 11, 12, 13, 15, 17, 18]              # utils/eval.py:11
synthia_set_16_to_13 = [0, 1, 2, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]         # utils/eval.py:12


def _to_device_labels(a, device, num_class, is_gt):
    """numpy / tensor label map -> contiguous int64 CUDA tensor with the reference's
    conversions: float ground truth is range-checked as float, then truncated
    (``gt_image[mask].astype('int')``, utils/eval.py:111-112)."""
    if isinstance(a, torch.Tensor) and a.dtype == torch.int64 and a.is_cuda and a.device == device:
        return a if a.is_contiguous() else a.contiguous()           # what a CUDA caller passes: nothing to convert
    t = torch.as_tensor(a) if not isinstance(a, torch.Tensor) else a
    if t.device != device:
        if t.device.type == "cpu" and t.numel() > 0:
            t = t.contiguous()
            try:
                t = t.pin_memory()
            except RuntimeError:
                pass
        t = t.to(device, non_blocking=True)
    if t.is_floating_point():
        if is_gt:
            ok = (t >= 0) & (t < num_class)
            t = torch.where(ok, t.trunc().to(torch.int64), torch.full((), -1, dtype=torch.int64, device=device))
        else:
            t = t.to(torch.int64)
    elif t.dtype != torch.int64:
        t = t.to(torch.int64)
    return t.contiguous()


class Eval:
    """``Eval(num_class)`` of ``utils/eval.py``.  Extra, all optional:

    :param device: CUDA device of the accumulator (default: the current one)
    :param defer:  K > 0: ``add_batch`` calls with CUDA tensors are QUEUED and run K at a time in ONE launch
                   (``msq_confusion_i64_multi``): the reference's validation loops call ``add_batch`` once per image
                   (``tools/train_source.py:429-492``), and one 8 MB launch per image is launch-bound.  The queue is
                   flushed when it holds K pairs and whenever the matrix or a metric is read.  The queued tensors
                   must not be modified in place before that (checked: a changed ``_version`` raises).
    """

    def __init__(self, num_class, device=None, defer=0):
        if not torch.cuda.is_available():
            raise RuntimeError("maxsquareloss_b200.Eval needs a CUDA device: there is no CPU fallback")
        if num_class < 1 or num_class > _lib.MAX_CLASSES:
            raise RuntimeError(f"num_class must be in [1, {_lib.MAX_CLASSES}]")
        _lib.load()
        self.num_class = num_class
        self.ignore_index = None
        self.synthia = True if num_class == 16 else False
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self._host = np.zeros((num_class,) * 2)
        # [C*C counts | 2 error flags (as one int64: two uint32)]
        self._dev = torch.zeros(num_class * num_class + 1, dtype=torch.int64, device=self.device)
        self._pending = False
        self._dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self._cm_ptr = self._dev.data_ptr()
        self._err_ptr = self._cm_ptr + 8 * num_class * num_class
        self._call = _lib.load().msq_confusion_i64
        self.defer = int(defer)
        self._queue = []                 # deferred (gt, pred, gt._version, pred._version)
        self._per_image = []             # device blocks (k, C*C) int64 of add_batch_per_image, in call order
        # the deferred queue's pointer table is filled as the calls come in, so that a flush is ONE library call
        kq = max(self.defer, 1)
        self._q_gt, self._q_pr = (ctypes.c_void_p * kq)(), (ctypes.c_void_p * kq)()
        self._q_np = (ctypes.c_int64 * kq)()
        self._multi = _lib.load().msq_confusion_i64_multi

    # ------------------------------------------------------------------ accumulate
    def add_batch(self, gt_image, pre_image):
        """``gt_image``: (N,H,W) labels, -1 / >= C ignored.  ``pre_image``: (N,H,W) class
        ids, or floating CUDA logits (N,C,H,W) to have the argmax taken on the device."""
        if isinstance(pre_image, torch.Tensor) and pre_image.is_floating_point() and \
                pre_image.dim() == np.ndim(gt_image) + 1:
            return self.add_batch_logits(gt_image, pre_image)
        # assert the size of two images are same (utils/eval.py:119)
        assert gt_image.shape == pre_image.shape
        host_call = not (isinstance(gt_image, torch.Tensor) and gt_image.is_cuda and
                         isinstance(pre_image, torch.Tensor) and pre_image.is_cuda)
        gt = _to_device_labels(gt_image, self.device, self.num_class, True)
        pr = _to_device_labels(pre_image, self.device, self.num_class, False)
        if self.defer > 0 and not host_call:
            q = self._queue
            k = len(q)
            self._q_gt[k], self._q_pr[k], self._q_np[k] = gt.data_ptr(), pr.data_ptr(), gt.numel()
            q.append((gt, pr, gt._version, pr._version))
            if k + 1 >= self.defer:
                self._flush_queue()
            return
        rc = self._call(gt.data_ptr(), pr.data_ptr(), gt.numel(), self.num_class, self._cm_ptr, self._err_ptr,
                        _raw_stream(self._dev_index))
        if rc:
            _lib.check(rc)
        self._pending = True
        if host_call:
            # numpy callers get the reference's synchronous behaviour, incl. ValueError now
            self._fold()

    def _launch_multi(self, pairs, cm_ptr, cm_stride, total_ptr):
        """One launch for ``pairs`` = [(gt, pred), ...] (C ABI ``msq_confusion_i64_multi``)."""
        k = len(pairs)
        arr = ctypes.c_void_p * k
        gts = arr(*[p[0].data_ptr() for p in pairs])
        prs = arr(*[p[1].data_ptr() for p in pairs])
        npx = (ctypes.c_int64 * k)(*[p[0].numel() for p in pairs])
        _lib.check(_lib.load().msq_confusion_i64_multi(gts, prs, npx, k, self.num_class, cm_ptr, cm_stride, total_ptr,
                                                       self._err_ptr, _raw_stream(self._dev_index)))
        self._pending = True

    def _flush_queue(self):
        if not self._queue:
            return
        q, self._queue = self._queue, []
        for gt, pr, vg, vp in q:
            if gt._version != vg or pr._version != vp:
                raise RuntimeError("Eval(defer=K): a tensor handed to add_batch was modified in place before the queued "
                                   "launch ran; pass fresh tensors or read a metric / call flush() first")
        rc = self._multi(self._q_gt, self._q_pr, self._q_np, len(q), self.num_class, self._cm_ptr, 0, None, self._err_ptr,
                         _raw_stream(self._dev_index))            # the table was filled as the calls came in
        if rc:
            _lib.check(rc)
        self._pending = True

    def flush(self):
        """Run the queued ``add_batch`` calls now (``defer`` > 0); no host synchronisation."""
        self._flush_queue()

    def add_batch_per_image(self, gt_image, pre_image):
        """``add_batch`` that ALSO keeps one confusion matrix per image of the batch: what ``tools/analysis.py:200-229``
        obtains with ``Eval.add_batch`` / metrics / ``Eval.reset()`` per image next to ``totalEval.add_batch``, in ONE
        launch for the whole batch and without a host round trip per image.  ``pre_image``: (N,H,W) class ids or
        float32 CUDA logits (N,C,H,W).  Read the results with ``per_image_matrices()`` / ``per_image_metrics()``."""
        c2 = self.num_class ** 2
        self._flush_queue()
        if isinstance(pre_image, torch.Tensor) and pre_image.is_floating_point() and pre_image.dim() == np.ndim(gt_image) + 1:
            if not (pre_image.is_cuda and pre_image.dtype == torch.float32):
                raise RuntimeError("logits must be a float32 CUDA tensor (N,C,H,W)")
            n, c = pre_image.shape[0], pre_image.shape[1]
            assert tuple(gt_image.shape) == (n,) + tuple(pre_image.shape[2:])
            if c != self.num_class:
                raise ValueError(f"logits have {c} classes, Eval was built with {self.num_class}")
            gt = _to_device_labels(gt_image, self.device, self.num_class, True)
            lg = pre_image.contiguous()
            block = torch.zeros(n, c2, dtype=torch.int64, device=self.device)
            _lib.check(_lib.load().msq_confusion_per_image_logits_f32(
                gt.data_ptr(), lg.data_ptr(), n, c, lg.numel() // max(n * c, 1), block.data_ptr(), self._cm_ptr,
                _raw_stream(self._dev_index)))
            self._pending = True
        else:
            assert tuple(gt_image.shape) == tuple(pre_image.shape)          # utils/eval.py:119
            gt = _to_device_labels(gt_image, self.device, self.num_class, True)
            pr = _to_device_labels(pre_image, self.device, self.num_class, False)
            n = gt.shape[0] if gt.dim() == 3 else 1
            g2, p2 = gt.reshape(n, -1), pr.reshape(n, -1)
            block = torch.zeros(n, c2, dtype=torch.int64, device=self.device)
            if n and g2.shape[1]:
                self._launch_multi([(g2[i], p2[i]) for i in range(n)], block.data_ptr(), c2, self._cm_ptr)
        self._per_image.append(block)

    def per_image_matrices(self):
        """(K,C,C) int64 ndarray: the confusion matrix of every image given to ``add_batch_per_image`` since the last
        ``reset()``, in order (one D2H of K*C*C*8 bytes; synchronises)."""
        self._fold()                       # surfaces the error flags of those launches
        if not self._per_image:
            return np.zeros((0, self.num_class, self.num_class), dtype=np.int64)
        return torch.cat(self._per_image).cpu().numpy().reshape(-1, self.num_class, self.num_class)

    def per_image_metrics(self):
        """[(PA, MPA, MIoU, FWIoU), ...] per image, each value what the reference's ``Eval`` holding only that image
        returns (``tools/analysis.py:177-183,212-214``): the same NumPy expressions on the float64 matrix."""
        res = []
        for m in self.per_image_matrices():
            h = HostEval(self.num_class, m)
            res.append((h.Pixel_Accuracy(), h.Mean_Pixel_Accuracy(), h.Mean_Intersection_over_Union(),
                        h.Frequency_Weighted_Intersection_over_Union()))
        return res

    def add_batch_logits(self, gt_image, logits):
        """Fused ``np.argmax(logits, axis=1)`` + ``add_batch`` (tools/train_source.py:457-459,492)."""
        if not (isinstance(logits, torch.Tensor) and logits.is_cuda and logits.dtype == torch.float32):
            raise RuntimeError("logits must be a float32 CUDA tensor (N,C,H,W)")
        n, c = logits.shape[0], logits.shape[1]
        assert tuple(gt_image.shape) == (n,) + tuple(logits.shape[2:])
        if c != self.num_class:
            raise ValueError(f"logits have {c} classes, Eval was built with {self.num_class}")
        gt = _to_device_labels(gt_image, self.device, self.num_class, True)
        lg = logits.contiguous()
        hw = lg.numel() // max(n * c, 1)
        self._flush_queue()
        _lib.check(_lib.load().msq_confusion_logits_f32(gt.data_ptr(), lg.data_ptr(), n, c, hw,
                                                        self._cm_ptr, _raw_stream(self._dev_index)))
        self._pending = True

    def add_batch_flip(self, gt_image, logits, logits_flipped):
        """Flip-ensemble evaluation (``tools/evaluate.py:120-141`` with ``--flip``), fused on the device:
        ``argmax((softmax(logits) + flip(softmax(logits_flipped), -1)) / 2)`` against ``gt_image``.
        ``logits_flipped`` is the model's output for the horizontally flipped input, not flipped back."""
        for t, name in ((logits, "logits"), (logits_flipped, "logits_flipped")):
            if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.dim() == 4):
                raise RuntimeError(f"{name} must be a float32 CUDA tensor (N,C,H,W)")
        assert tuple(logits.shape) == tuple(logits_flipped.shape)
        n, c, h, w = logits.shape
        assert tuple(gt_image.shape) == (n, h, w)
        if c != self.num_class:
            raise ValueError(f"logits have {c} classes, Eval was built with {self.num_class}")
        gt = _to_device_labels(gt_image, self.device, self.num_class, True)
        la, lb = logits.contiguous(), logits_flipped.contiguous()
        self._flush_queue()
        _lib.check(_lib.load().msq_confusion_flip_f32(gt.data_ptr(), la.data_ptr(), lb.data_ptr(), n, c, h, w,
                                                      self._cm_ptr, _raw_stream(self._dev_index)))
        self._pending = True

    def _fold(self):
        """device counts -> host float64 matrix (one small D2H, synchronises)."""
        self._flush_queue()
        if not self._pending:
            return
        c2 = self.num_class ** 2
        host = self._dev.cpu().numpy()
        self._dev.zero_()
        self._pending = False
        # the counts of every in-contract pixel are kept even when a batch tripped an error flag: the reference raises
        # inside the offending add_batch and leaves what was accumulated before it intact (utils/eval.py:113-121)
        self._host += host[:c2].reshape(self.num_class, self.num_class)
        flags = int(host[c2])
        if flags & 0xFFFFFFFF:
            raise ValueError("'list' argument must have no negative elements")      # numpy.bincount's message
        if flags >> 32:
            raise ValueError(f"cannot reshape array into shape ({self.num_class},{self.num_class})")

    @property
    def confusion_matrix(self):
        self._fold()
        return self._host

    @confusion_matrix.setter
    def confusion_matrix(self, value):
        self._queue = []
        self._dev.zero_()
        self._pending = False
        self._host = np.asarray(value, dtype=np.float64)

    def device_counts(self):
        self._flush_queue()
        return self._device_counts()

    def _device_counts(self):
        """(C,C) int64 CUDA tensor of the counts not yet folded to the host (for an
        NCCL all-reduce without a host round trip)."""
        return self._dev[:self.num_class ** 2].view(self.num_class, self.num_class)

    def reset(self):
        self._queue = []
        self._per_image = []
        self._dev.zero_()
        self._pending = False
        self._host = np.zeros((self.num_class,) * 2)

    # ------------------------------------------------------------------ metrics (utils/eval.py:22-106)
    def _pick(self, per_class, out_16_13):
        if self.synthia:
            return np.nanmean(per_class[:self.ignore_index]), np.nanmean(per_class[synthia_set_16_to_13])
        if out_16_13:
            return np.nanmean(per_class[synthia_set_16]), np.nanmean(per_class[synthia_set_13])
        return np.nanmean(per_class[:self.ignore_index])

    def _ratios(self):
        cm = self.confusion_matrix
        tp, rows, cols = np.diag(cm), cm.sum(axis=1), cm.sum(axis=0)
        with np.errstate(divide='ignore', invalid='ignore'):
            return cm, tp, rows, cols, tp / rows, tp / (rows + cols - tp), tp / cols

    def Pixel_Accuracy(self):
        cm = self.confusion_matrix
        if np.sum(cm) == 0:
            print("Attention: pixel_total is zero!!!")
            return 0
        return np.diag(cm).sum() / cm.sum()

    def Mean_Pixel_Accuracy(self, out_16_13=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return self._pick(self._ratios()[4], out_16_13)

    def Mean_Intersection_over_Union(self, out_16_13=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return self._pick(self._ratios()[5], out_16_13)

    def Mean_Precision(self, out_16_13=False):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return self._pick(self._ratios()[6], out_16_13)

    def Frequency_Weighted_Intersection_over_Union(self, out_16_13=False):
        cm, tp, rows, cols, _, _, _ = self._ratios()
        with np.errstate(divide='ignore', invalid='ignore'):
            fw = np.multiply(rows, tp) / (rows + cols - tp)
        total = np.sum(cm)

        def nansum_in_order(v):     # the reference sums the non-NaN entries left to right
            s = 0
            for x in v:
                if not np.isnan(x):
                    s = s + x
            return s / total
        with np.errstate(divide='ignore', invalid='ignore'):
            if self.synthia:
                return nansum_in_order(fw), nansum_in_order(fw[synthia_set_16_to_13])
            if out_16_13:
                return nansum_in_order(fw[synthia_set_16]), nansum_in_order(fw[synthia_set_13])
            return nansum_in_order(fw)

    def Print_Every_class_Eval(self, out_16_13=False):
        cm, _, rows, cols, mpa, miou, prec = self._ratios()
        with np.errstate(divide='ignore', invalid='ignore'):
            class_ratio = rows / np.sum(cm)
            pred_ratio = cols / np.sum(cm)
        print('===>Everyclass:\t' + 'MPA\t' + 'MIoU\t' + 'PC\t' + 'Ratio\t' + 'Pred_Retio')
        if out_16_13:
            miou = miou[synthia_set_16]

        def pct(v):
            return str(round(v * 100, 2)) if not np.isnan(v) else 'nan'
        for k in range(len(miou)):
            print('===>' + name_classes[k] + ':\t' + pct(mpa[k]) + '\t' + pct(miou[k]) + '\t' + pct(prec[k]) +
                  '\t' + pct(class_ratio[k]) + '\t' + pct(pred_ratio[k]))


class HostEval(Eval):
    """The metric methods of ``Eval`` on a given (C,C) matrix; no device, no accumulation (per-image metrics)."""

    def __init__(self, num_class, matrix):          # noqa: super().__init__ needs a GPU and is not wanted here
        self.num_class = num_class
        self.ignore_index = None
        self.synthia = True if num_class == 16 else False
        self._host = np.asarray(matrix, dtype=np.float64)

    @property
    def confusion_matrix(self):
        return self._host


def fast_hist(gt, pred, num_class, device=None):
    """(C,C) int64 confusion matrix of one (gt, pred) pair; rows = ground truth.
    Public alias of the reference's private ``Eval._Eval__generate_matrix``
    (``utils/eval.py:109-115``).  ``pred`` may be class ids or CUDA logits."""
    ev = Eval(num_class, device=device)
    ev.add_batch(gt, pred)
    return ev.confusion_matrix.astype(np.int64)
