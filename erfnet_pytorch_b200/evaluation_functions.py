"""Losses of train/evaluation_functions.py on the B200 kernels.  ``CrossEntropyLoss2d`` keeps the reference's
constructor and call signature (train/evaluation_functions.py:6-14); unlike the reference it HONOURS ``weight``
(the reference silently drops it -- SURVEY D3); ``weight=None`` reproduces the reference exactly."""
from __future__ import annotations

import torch

from . import ops


class _CEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, weight, ignore_index, c, nchw):
        lsum, wsum, dlog = ops.cross_entropy(logits, target, weight, ignore_index, c, nchw, want_grad=logits.requires_grad)
        ctx.dlog = dlog
        return (lsum / wsum).reshape(())

    @staticmethod
    def backward(ctx, go):
        d = ctx.dlog          # kept on the node: a second backward through a retained graph reuses it
        return d * go.to(d.dtype), None, None, None, None, None


class CrossEntropyLoss2d(torch.nn.Module):
    """loss = sum_valid w[y] * -log_softmax(outputs)[y] / sum_valid w[y], ignore_index = -1.
    ``outputs``: NCHW fp32 logits (the reference API) -- or the model's internal NHWC logits
    (``ERFNet.forward_nhwc``) with ``nhwc_classes=C``, which skips the NCHW fp32 round trip."""

    def __init__(self, weight=None, ignore_index=-1, strict_labels=False):
        """strict_labels: the kernels skip EVERY label outside [0, C) (that is how ignore_index is honoured), whereas torch's
        nn.CrossEntropyLoss asserts on an out-of-range label that is not ignore_index (an un-relabelled 255, an off-by-one class id).
        strict_labels=True restores that behaviour with a host check of target.min() / max() (one device synchronisation per call:
        meant for the first steps of a run / debugging a data pipeline, not for the steady state)."""
        super().__init__()
        self.register_buffer("weight", None if weight is None else torch.as_tensor(weight, dtype=torch.float32))
        self.ignore_index = ignore_index
        self.strict_labels = strict_labels

    def _check_labels(self, targets, c):
        if targets.numel() == 0:
            return
        bad = (targets != self.ignore_index) & ((targets < 0) | (targets >= c))
        if bool(bad.any()):
            v = int(targets[bad][0])
            raise RuntimeError(f"CrossEntropyLoss2d: target {v} is out of bounds for {c} classes (ignore_index = {self.ignore_index})")

    def forward(self, outputs, targets, nhwc_classes=None):
        w = None if self.weight is None else self.weight.to(outputs.device)
        if self.strict_labels:
            self._check_labels(targets, nhwc_classes if nhwc_classes is not None else outputs.shape[1])
        if nhwc_classes is not None:
            return _CEFn.apply(outputs, targets, w, self.ignore_index, nhwc_classes, False)
        return _CEFn.apply(outputs.contiguous(), targets, w, self.ignore_index, outputs.shape[1], True)


# ---------------------------------------------------------------------------------------------------------------
# Masked regression losses of the fork (train/evaluation_functions.py:32-131): one reduction kernel forward, one
# gradient kernel backward; the "labelled pixel" mask (target != -1) is evaluated inside the kernels.
class _MaskedLossFn(torch.autograd.Function):
    """loss = scale * sum_valid(per-pixel loss) / #valid.  Operands: a, b (per-pixel or per-class), prob ([B,C,H,W] or None)."""

    @staticmethod
    def forward(ctx, kind, a, b, prob, target, eps, scale):
        af, bf, pf = ops._f32c(a, "prediction"), ops._f32c(b, "variance"), ops._f32c(prob, "class probabilities")
        tf = ops._f32c(target, "target")
        if tf.numel() == 0:
            raise RuntimeError("masked loss: empty target")
        ops.check_masked_loss_shapes(kind, af, bf, pf, tf)
        acc = torch.zeros(2, device=tf.device, dtype=torch.float64)
        ops.masked_loss_fwd(kind, af, bf, pf, tf, eps, acc)
        ops.loss_global_reduce(acc)              # data-parallel training: (sum, #labelled pixels) over the global batch
        ctx.kind, ctx.eps, ctx.scale = kind, eps, scale
        ctx.shapes = (a.shape, None if b is None else b.shape)
        ctx.save_for_backward(af, bf, pf, tf, acc)
        return (acc[0] * scale / acc[1]).float()

    @staticmethod
    def backward(ctx, go):
        af, bf, pf, tf, acc = ctx.saved_tensors
        go = go.reshape(1).float().contiguous()
        r0, r1 = ops.masked_loss_bwd(ctx.kind, af, bf, pf, tf, ctx.eps, ctx.scale, acc, go)
        sa, sb = ctx.shapes
        if pf is None:
            da = r0.reshape(sa)
            db = None if r1 is None else r1.reshape(sb)
            dt = -da.reshape(tf.shape) if (ctx.kind == ops.LOSS_MSE_WEIGHTED and ctx.needs_input_grad[4]) else None
            return None, da, db, None, dt, None, None
        c = pf.shape[1]
        g = go.double() * ctx.scale / acc[1]
        da = (r1[:c] * g).float().reshape(sa)
        db = None if bf is None else (r1[c:] * g).float().reshape(sb)
        return None, da, db, r0, None, None, None


class L1LossMasked(torch.nn.Module):
    """mean over labelled pixels of |outputs - targets| -- train/evaluation_functions.py:89-97.
    outputs [B,1,H,W] (or [B,H,W]), targets float [B,H,W] with -1 = unlabelled."""

    def forward(self, outputs, targets):
        return _MaskedLossFn.apply(ops.LOSS_L1, outputs.squeeze(1), None, None, targets, 0.0, 1.0)


class LogLikelihoodLossMasked(torch.nn.Module):
    """mean over labelled pixels of log(2 pi (var+eps)) + (cost-t)^2/(var+eps) -- train/evaluation_functions.py:71-87."""

    def __init__(self, opt_eps=0.0):
        super().__init__()
        self.opt_eps = opt_eps

    def forward(self, output_cost, output_var, targets):
        return _MaskedLossFn.apply(ops.LOSS_LOGLIK, output_cost, output_var, None, targets, float(self.opt_eps), 1.0)


class L1LossClassProbMasked(torch.nn.Module):
    """mean over labelled pixels of sum_c |cost_c - t| prob_c -- train/evaluation_functions.py:32-47.
    output_prob [B,C,H,W], output_cost [1,C,1,1] (the net's class_power), targets [B,H,W]."""

    def forward(self, output_prob, output_cost, targets):
        return _MaskedLossFn.apply(ops.LOSS_L1_CLASSPROB, _per_class(output_cost, output_prob), None, output_prob, targets, 0.0, 1.0)


class LogLikelihoodLossClassProbMasked(torch.nn.Module):
    """mean over labelled pixels of sum_c (log(2 pi v_c) + (cost_c - t)^2 / v_c) prob_c, v_c = var_c + eps -- l.49-69."""

    def __init__(self, opt_eps=0.0):
        super().__init__()
        self.opt_eps = opt_eps

    def forward(self, output_prob, output_cost, output_var, targets):
        return _MaskedLossFn.apply(ops.LOSS_LOGLIK_CLASSPROB, _per_class(output_cost, output_prob), _per_class(output_var, output_prob),
                                   output_prob, targets, float(self.opt_eps), 1.0)


class MSELossWeighted(torch.nn.Module):
    """((outputs - targets)^2 * scale_factor * weight).mean() with a scalar weight -- train/evaluation_functions.py:122-131
    (the mean-teacher consistency loss, train/main.py:191, 341, 353)."""

    def __init__(self, scale_factor):
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, outputs, targets, weight):
        if outputs.shape != targets.shape:
            raise RuntimeError(f"MSELossWeighted: shape mismatch {tuple(outputs.shape)} vs {tuple(targets.shape)}")
        return _MaskedLossFn.apply(ops.LOSS_MSE_WEIGHTED, outputs, None, None, targets, 0.0, float(self.scale_factor) * float(weight))


def _per_class(t, prob):
    if t.numel() != prob.shape[1]:
        raise RuntimeError(f"class-probability loss: expected {prob.shape[1]} per-class values, got shape {tuple(t.shape)}")
    return t


# ---------------------------------------------------------------------------------------------------------------
# Accuracy metrics (train/evaluation_functions.py:133-163) from the iouEval confusion histogram: one argmax kernel and one
# histogram kernel instead of 2C+3 reductions.  Labels must lie in [-1, C) (-1 = unlabelled), as the fork's do.
def _confusion(prob, targets):
    c = prob.shape[1]
    pred = ops.argmax(prob.contiguous(), c, True)
    cm = torch.zeros(c, c, device=prob.device, dtype=torch.int64)
    ops.confusion_hist(pred, targets.to(torch.int64), c, cm)
    return cm


class ClassificationAccuracy:
    """#(argmax == target) / #(target >= 0) -- train/evaluation_functions.py:133-139."""

    def __call__(self, prob, targets):
        cm = _confusion(prob, targets)
        return cm.diagonal().sum().float() / cm.sum().float()


class MeanAccuracy:
    """Per-class accuracy CM[c,c] / sum_p CM[c,p] (NaN for absent classes) and its mean over present classes -- l.141-163."""

    def __init__(self, n_classes):
        self.n_classes = n_classes

    def __call__(self, prob, targets):
        cm = _confusion(prob, targets)[: self.n_classes]
        correct, total = cm.diagonal().float(), cm.sum(1).float()
        class_acc = correct / total
        present = total > 0
        mean_acc = torch.where(present, class_acc, torch.zeros_like(class_acc)).sum() / present.sum()
        return mean_acc, class_acc.cpu().numpy()
