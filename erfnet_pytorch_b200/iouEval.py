"""iouEval of eval/iouEval.py on a bit-exact integer confusion histogram.

Same constructor, ``addBatch(x, y)``, ``getIoU()``, ``reset()`` and float64 ``tp/fp/fn`` attributes.  The
reference materialises two fp32 one-hot tensors and nine reductions per batch; here one kernel accumulates the
nClasses x nClasses confusion matrix CM[g, p] in u64 and tp/fp/fn are read off it:
tp = diag, fp[c] = sum_{g != ignore} CM[g, c] - tp, fn[c] = sum_p CM[c, p] - tp.
"""
from __future__ import annotations

import torch

from . import ops


class iouEval:
    def __init__(self, nClasses, ignoreIndex=19):
        self.nClasses = nClasses
        self.ignoreIndex = ignoreIndex if nClasses > ignoreIndex else -1  # eval/iouEval.py:12
        self.reset()

    def reset(self):
        classes = self.nClasses if self.ignoreIndex == -1 else self.nClasses - 1
        self._classes = classes
        self._cm = None  # device u64 (as int64) [nClasses, nClasses], allocated on first addBatch
        self._tp = torch.zeros(classes).double()
        self._fp = torch.zeros(classes).double()
        self._fn = torch.zeros(classes).double()
        self._dirty = False

    # tp / fp / fn are attributes of the reference class (eval/iouEval.py:15-19): kept readable at any time, synchronised lazily
    @property
    def tp(self):
        self._sync()
        return self._tp

    @property
    def fp(self):
        self._sync()
        return self._fp

    @property
    def fn(self):
        self._sync()
        return self._fn

    def addBatch(self, x, y):  # x = preds, y = targets; "batch_size x 1 x H x W" label maps (or [B,H,W])
        if not (x.is_cuda and y.is_cuda) and not ops.EMULATED:
            if not torch.cuda.is_available():
                raise RuntimeError("iouEval.addBatch needs CUDA tensors (no CPU fallback)")
            x, y = x.cuda(), y.cuda()
        # eval/iouEval.py:37-46: a [B,C,H,W] input with C > 1 is "already one-hot"; for 0/1 one-hot maps the reference's products
        # sum(x_onehot * y_onehot) are exactly the histogram of the arg-max label maps, which one kernel recovers
        if x.dim() == 4 and x.size(1) != 1:
            x = ops.argmax(x.float().contiguous(), x.size(1), True).unsqueeze(1)
        if y.dim() == 4 and y.size(1) != 1:
            y = ops.argmax(y.float().contiguous(), y.size(1), True).unsqueeze(1)
        if x.dtype != y.dtype:
            x, y = x.long(), y.long()
        if self._cm is None or self._cm.device != x.device:
            self._cm = torch.zeros(self.nClasses, self.nClasses, device=x.device, dtype=torch.int64)
        ops.confusion_hist(x, y, self.nClasses, self._cm)
        self._dirty = True

    def _sync(self):
        if self._dirty:
            kept = self._classes if self.ignoreIndex == -1 else min(self._classes, self.ignoreIndex)
            t = ops.iou_counts(self._cm, self.nClasses, self.ignoreIndex, kept).cpu()
            self._tp = torch.zeros(self._classes).double()
            self._fp = torch.zeros(self._classes).double()
            self._fn = torch.zeros(self._classes).double()
            self._tp[:kept], self._fp[:kept], self._fn[:kept] = t[0], t[1], t[2]
            self._dirty = False

    def confusion_matrix(self):
        return None if self._cm is None else self._cm.clone()

    def merge_(self, cm):
        """Add another rank's confusion matrix (after an all-reduce)."""
        self._cm = cm.clone() if self._cm is None else self._cm + cm
        self._dirty = True

    def getIoU(self):
        self._sync()
        num = self._tp
        den = self._tp + self._fp + self._fn + 1e-15
        iou = num / den
        return torch.mean(iou), iou  # "iou mean", "iou per class"
