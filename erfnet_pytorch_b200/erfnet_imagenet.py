"""The ImageNet pre-training net (imagenet/erfnet_imagenet.py, BASELINE config 3) on the B200 kernels: same class names, constructor
signatures and attribute names (= checkpoint keys ``features.encoder.*``, ``classifier.linear.*``).  The encoder is the Cityscapes encoder
without ``output_conv`` (erfnet_imagenet.py:60-86); the tail -- MaxPool2d(2) -> AvgPool2d(14) -> view -> Linear(128, num_classes)
(l.89-110) -- runs on the pooling / small-GEMM kernels of csrc/tail_io.cu.  ``CrossEntropyLoss`` / ``accuracy`` / ``FlatSGD`` /
``adjust_learning_rate`` restate imagenet/main.py:80-85, 168-189, 276-301 on the row-wise CE and flat SGD kernels."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import erfnet as _E
from . import ops


class DownsamplerBlock(_E.DownsamplerBlock):
    """imagenet/erfnet_imagenet.py:11-22."""

    def __init__(self, ninput, noutput):
        super().__init__(ninput, noutput)


class non_bottleneck_1d(_E.non_bottleneck_1d):
    """imagenet/erfnet_imagenet.py:24-57."""

    def __init__(self, chann, dropprob, dilated):
        super().__init__(chann, dropprob, dilated)
        for name in ("bn1", "bn2", "dropout"):      # this file registers all four convs before the BatchNorms (checkpoint key order)
            self._modules[name] = self._modules.pop(name)


class Encoder(_E._B200Module):
    """imagenet/erfnet_imagenet.py:60-86: the ERFNet encoder without the 1x1 prediction conv."""

    def __init__(self):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(16, 64))
        for _ in range(0, 5):
            self.layers.append(non_bottleneck_1d(64, 0.1, 1))
        self.layers.append(DownsamplerBlock(64, 128))
        for _ in range(0, 2):
            for d in (2, 4, 8, 16):
                self.layers.append(non_bottleneck_1d(128, 0.1, d))
        self._in_ld, self._out_c = 4, 128

    def _nhwc(self, x, ctx):
        if x.shape[1] % 8 or x.shape[2] % 8:
            raise RuntimeError(f"Encoder: H and W must be multiples of 8, got {x.shape[1]}x{x.shape[2]}")
        y = _E._run(self.initial_block, x, ctx)
        for layer in self.layers:
            y = _E._run(layer, y, ctx)
        return y


class _MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2, stride 2) on an NHWC tensor (all its channels)."""

    @staticmethod
    def forward(ctx, x):
        n, h, w, c = x.shape
        out = torch.empty(n, h // 2, w // 2, c, device=x.device, dtype=x.dtype)
        ops.maxpool_fwd(x, c, out, 0)
        ctx.save_for_backward(x)
        return out

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dx = torch.zeros_like(x)
        ops.maxpool_bwd(x, x.shape[-1], dy.contiguous(), 0, dx)
        return dx


class _AvgPoolFn(torch.autograd.Function):
    """nn.AvgPool2d(k, 1, 0): NHWC activations in, NCHW fp32 out (the boundary tensor the reference hands to Classifier)."""

    @staticmethod
    def forward(ctx, x, k):
        ctx.k, ctx.shape, ctx.dtype = k, x.shape, x.dtype
        return ops.avgpool_fwd(x, x.shape[-1], k).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dy):
        n, h, w, ld = ctx.shape
        return ops.avgpool_bwd(dy.permute(0, 2, 3, 1).contiguous(), h, w, ld, ctx.k, ctx.dtype), None


class Features(_E._B200Module):
    """imagenet/erfnet_imagenet.py:89-102."""

    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.extralayer1 = nn.MaxPool2d(2, stride=2)
        self.extralayer2 = nn.AvgPool2d(14, 1, 0)

    def forward(self, input):
        return self._pooled(input, _E._Ctx(self._precision(), self.training))

    def _pooled(self, input, ctx):
        y = self.encoder._nhwc(_E.input_to_nhwc(input, 4, ctx.dtype), ctx)
        y = _MaxPool2Fn.apply(y)
        if y.shape[1] < 14 or y.shape[2] < 14:
            raise RuntimeError(f"Features: AvgPool2d(14) needs at least 14x14 after the pooling, got {y.shape[1]}x{y.shape[2]}")
        return _AvgPoolFn.apply(y, 14)


class _LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        x = x.contiguous()
        ctx.save_for_backward(x)
        ctx.w, ctx.b = w, b
        return ops.linear_fwd(x, w.detach(), b.detach())

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        w, b = ctx.w, ctx.b
        dx, dw, db = ops.linear_bwd(x, w.detach(), dy, ctx.needs_input_grad[0], _E._sink(w), _E._sink(b))
        grads = (_E._ret(w, dw), _E._ret(b, db))
        _E._sunk(w, b)
        return (dx,) + grads


class Classifier(_E._B200Module):
    """imagenet/erfnet_imagenet.py:104-112."""

    def __init__(self, num_classes):
        super().__init__()
        self.linear = nn.Linear(128, num_classes)

    def forward(self, input):
        output = input.reshape(input.size(0), 128)     # first is batch_size; fails like the reference's view() unless 128 x 1 x 1
        return _LinearFn.apply(output.float(), self.linear.weight, self.linear.bias)


class ERFNet(_E._B200Module):
    """imagenet/erfnet_imagenet.py:114-125."""

    def __init__(self, num_classes):
        super().__init__()
        self.features = Features()
        self.classifier = Classifier(num_classes)

    def forward(self, input):
        ctx = _E._Ctx(self._precision(), self.training)      # the net's precision governs the whole pass
        return self.classifier(self.features._pooled(input, ctx))


# ---- imagenet/main.py pieces ---------------------------------------------------------------------------------------------------------
class _CERowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, target, ignore_index):
        logits = logits.contiguous()
        acc, lse, rank = ops.ce_rows_fwd(logits, target.contiguous(), ignore_index)
        ops.loss_global_reduce(acc)              # data-parallel training: (sum of -log p, #rows) over the global batch
        ctx.save_for_backward(logits, target, lse, acc)
        ctx.ignore_index = ignore_index
        ctx.mark_non_differentiable(rank)
        return (acc[0] / acc[1]).reshape(()), rank

    @staticmethod
    def backward(ctx, go, _):
        logits, target, lse, acc = ctx.saved_tensors
        return ops.ce_rows_bwd(logits, target.contiguous(), ctx.ignore_index, lse, acc, go.contiguous().float().reshape(1)), None, None


class CrossEntropyLoss(nn.Module):
    """``nn.CrossEntropyLoss()`` of imagenet/main.py:80 on [N, C] logits: mean over the rows whose target is not ``ignore_index``.
    ``last_rank`` keeps the rank of every row's target logit from the same pass (what ``accuracy`` needs)."""

    def __init__(self, ignore_index=-100):
        super().__init__()
        self.ignore_index = ignore_index
        self.last_rank = None

    def forward(self, output, target):
        if output.dim() != 2 or output.dtype != torch.float32:
            raise RuntimeError(f"CrossEntropyLoss: expected fp32 [N, C] logits, got {output.dtype} {tuple(output.shape)}")
        loss, self.last_rank = _CERowsFn.apply(output, target, self.ignore_index)
        return loss


def accuracy(output, target, topk=(1,)):
    """imagenet/main.py:288-301: precision@k in percent for each k, from the per-row rank of the target logit (one kernel)."""
    with torch.no_grad():
        _, _, rank = ops.ce_rows_fwd(output.detach().contiguous().float(), target.contiguous(), -100)
        batch_size = target.size(0)
        return [((rank >= 0) & (rank < k)).sum().float().mul_(100.0 / batch_size).reshape(1) for k in topk]


def adjust_learning_rate(optimizer, epoch, base_lr):
    """imagenet/main.py:276-286: constant lr and weight decay 1e-4 up to epoch 15, then lr * 0.95^(epoch-15) and no weight decay."""
    lr, wd, milestone = base_lr, 1e-4, 15
    if epoch > milestone:
        lr = base_lr * (0.95 ** (epoch - milestone))
        wd = 0
    for param_group in optimizer.param_groups:
        param_group["lr"] = lr
        param_group["weight_decay"] = wd
