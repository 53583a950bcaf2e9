"""NumPy port of the reference's evaluation accumulator (``utils/eval.py``).
TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

* ``confusion``     <- ``Eval.__generate_matrix``  (``utils/eval.py:109-115``)
* ``EvalPort``      <- ``Eval``                    (``utils/eval.py:14-124``)

Rows are ground truth, columns are predictions.  ``gt`` values outside [0,C)
(-1, 255, ...) are ignored; float ``gt`` is truncated by ``astype('int')``;
a prediction equal to C silently aliases into the next row, a negative
flattened index raises ``ValueError`` (numpy.bincount), an index >= C*C makes
the reshape raise ``ValueError``.
"""
import numpy as np

SYNTHIA_16_OF_19 = [0, 1, 2, 3, 4, 5, 6, 7, 8, 10, 11, 12, 13, 15, 17, 18]   # utils/eval.py:10
SYNTHIA_13_OF_19 = [0, 1, 2, 6, 7, 8, 10, 11, 12, 13, 15, 17, 18]            # utils/eval.py:11
SYNTHIA_13_OF_16 = [0, 1, 2, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15]             # utils/eval.py:12


def confusion(gt: np.ndarray, pre: np.ndarray, num_class: int) -> np.ndarray:
    valid = (gt >= 0) & (gt < num_class)
    flat = num_class * gt[valid].astype('int') + pre[valid]
    counts = np.bincount(flat, minlength=num_class ** 2)
    return counts.reshape(num_class, num_class)


class EvalPort:
    def __init__(self, num_class: int):
        self.num_class = num_class
        self.synthia = num_class == 16
        self.reset()

    def reset(self):
        self.confusion_matrix = np.zeros((self.num_class, self.num_class))

    def add_batch(self, gt, pre):
        assert gt.shape == pre.shape
        self.confusion_matrix += confusion(gt, pre, self.num_class)

    # ---- metrics: each one is the reference expression on the float64 matrix ----
    def _subsets(self, v, out_16_13):
        if self.synthia:
            return np.nanmean(v), np.nanmean(v[SYNTHIA_13_OF_16])
        if out_16_13:
            return np.nanmean(v[SYNTHIA_16_OF_19]), np.nanmean(v[SYNTHIA_13_OF_19])
        return np.nanmean(v)

    def iou_per_class(self):
        cm = self.confusion_matrix
        with np.errstate(divide='ignore', invalid='ignore'):
            return np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))

    def Mean_Intersection_over_Union(self, out_16_13=False):     # utils/eval.py:45-59
        return self._subsets(self.iou_per_class(), out_16_13)

    def Pixel_Accuracy(self):                                    # utils/eval.py:22-29
        cm = self.confusion_matrix
        if cm.sum() == 0:
            return 0
        return np.diag(cm).sum() / cm.sum()

    def Mean_Pixel_Accuracy(self, out_16_13=False):              # utils/eval.py:31-43
        cm = self.confusion_matrix
        with np.errstate(divide='ignore', invalid='ignore'):
            return self._subsets(np.diag(cm) / cm.sum(axis=1), out_16_13)

    def Mean_Precision(self, out_16_13=False):                   # utils/eval.py:77-88
        cm = self.confusion_matrix
        with np.errstate(divide='ignore', invalid='ignore'):
            return self._subsets(np.diag(cm) / cm.sum(axis=0), out_16_13)

    def Frequency_Weighted_Intersection_over_Union(self, out_16_13=False):   # utils/eval.py:61-75
        cm = self.confusion_matrix
        with np.errstate(divide='ignore', invalid='ignore'):
            fw = cm.sum(axis=1) * np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))

        def tot(v):
            # the reference sums a generator with the builtin-style np.sum: plain
            # left-to-right float64 addition of the non-NaN entries
            s = 0
            for x in v:
                if not np.isnan(x):
                    s = s + x
            return s / cm.sum()
        if self.synthia:
            return tot(fw), tot(fw[SYNTHIA_13_OF_16])
        if out_16_13:
            return tot(fw[SYNTHIA_16_OF_19]), tot(fw[SYNTHIA_13_OF_19])
        return tot(fw)


def flip_ensemble_argmax(pred, pred_flip):
    """``tools/evaluate.py:120-136`` (``--flip``): softmax of the prediction and of the prediction for the
    horizontally flipped image, the latter flipped back, averaged; ``np.argmax`` over classes.
    ``pred`` / ``pred_flip``: torch tensors (N,C,H,W) as the model returns them.  -> numpy int64 (N,H,W)."""
    import torch
    import torch.nn.functional as F

    def flip(x, dim):                                   # evaluate.py:122-127
        dim = x.dim() + dim if dim < 0 else dim
        inds = tuple(slice(None, None) if i != dim
                     else x.new(torch.arange(x.size(i) - 1, -1, -1).tolist()).long()
                     for i in range(x.dim()))
        return x[inds]
    pred_P = F.softmax(pred, dim=1)
    pred_P_flip = F.softmax(pred_flip, dim=1)
    pred_P_2 = flip(pred_P_flip, -1)
    pred_c = (pred_P + pred_P_2) / 2
    return np.argmax(pred_c.data.cpu().numpy(), axis=1)
