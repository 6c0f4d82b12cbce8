"""The fork's block module (train/erfnet_blocks.py) on the B200 kernels: same class names and constructor signatures,
with the FORK's defaults -- BatchNorm off, dropout off unless ``use_dropout`` -- so that ``from erfnet_blocks import *``
in the fork's net_*.py files resolves to these.  The BatchNorm modules are still constructed (their parameters and
buffers are part of the fork's checkpoints) but take no part in the arithmetic when ``use_batchnorm`` is False.

Built: DownsamplerBlock, non_bottleneck_1d, UpsamplerBlock, DecoderBlockSmall, DecoderBlock, LadderBlock, SoftMaxConv (train +
eval, all on the same kernels as erfnet.py)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops

from .ops import ConvSpec

from . import erfnet as _E


class DownsamplerBlock(_E.DownsamplerBlock):
    """train/erfnet_blocks.py:15-29."""

    def __init__(self, ninput, noutput, use_batchnorm=False):
        super().__init__(ninput, noutput, use_batchnorm=use_batchnorm)


class non_bottleneck_1d(_E.non_bottleneck_1d):
    """train/erfnet_blocks.py:33-73."""

    def __init__(self, chann, dropprob, dilated, use_dropout=False, use_batchnorm=False):
        super().__init__(chann, dropprob, dilated, use_dropout=use_dropout, use_batchnorm=use_batchnorm)


class UpsamplerBlock(_E.UpsamplerBlock):
    """train/erfnet_blocks.py:77-88."""

    def __init__(self, ninput, noutput, use_batchnorm=False):
        super().__init__(ninput, noutput, use_batchnorm=use_batchnorm)


class DecoderBlockSmall(_E._B200Module):
    """train/erfnet_blocks.py:114-128: UpsamplerBlock -> non_bottleneck_1d(out, 0, 1); activations stay NHWC between the layers."""

    def __init__(self, in_channels, out_channels, use_dropout=False):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(UpsamplerBlock(in_channels, out_channels))
        self.layers.append(non_bottleneck_1d(out_channels, 0, 1, use_dropout=use_dropout))
        self._out_c = out_channels

    def _nhwc(self, x, ctx):
        for layer in self.layers:
            x = layer._nhwc(x, ctx)
        return x


class DecoderBlock(DecoderBlockSmall):
    """train/erfnet_blocks.py:132-135: one more non_bottleneck_1d."""

    def __init__(self, in_channels, out_channels, use_dropout=False):
        super().__init__(in_channels, out_channels, use_dropout)
        self.layers.append(non_bottleneck_1d(out_channels, 0, 1, use_dropout=use_dropout))


class LadderBlock(_E._B200Module):
    """train/erfnet_blocks.py:140-154: cat(input, input_ladder) -> non_bottleneck_1d(2C, 0.03, 1) -> Conv2d(2C, C, 3, padding 1) -> ReLU.
    The channel concat happens on the NHWC tensors (one copy), the 3x3 conv + bias + ReLU is one tcgen05 launch."""

    def __init__(self, noutput, use_batchnorm=False, use_dropout=False):
        super().__init__()
        if use_batchnorm:
            raise NotImplementedError("LadderBlock(use_batchnorm=True) is not used by any of the fork's nets and is not built")
        self.use_batchnorm = use_batchnorm
        self.layer = non_bottleneck_1d(2 * noutput, 0.03, 1, use_dropout=use_dropout)
        self.conv = nn.Conv2d(2 * noutput, noutput, 3, stride=1, padding=1, bias=True)
        self.bn = nn.BatchNorm2d(noutput, eps=_E.BN_EPS)
        self.noutput = noutput
        self._out_c = noutput

    def _nhwc2(self, x, ladder, ctx):
        y = torch.cat((x, ladder), dim=-1)                       # NHWC: channels are the last dimension
        y = self.layer._nhwc(y, ctx)
        return _E._run(_E._PlainConv(self.conv, ConvSpec(2 * self.noutput, self.noutput, 3, 3, 1, 1, 1), relu=True), y, ctx)

    def forward(self, input, input_ladder):
        ctx = _E._Ctx(self._precision(), self.training)
        x = _E._ToNHWC.apply(input, input.shape[1], ctx.dtype)
        l = _E._ToNHWC.apply(input_ladder, input_ladder.shape[1], ctx.dtype)
        return _E._ToNCHW.apply(self._nhwc2(x, l, ctx), self.noutput)


class _SoftmaxFn(torch.autograd.Function):
    """NHWC logits -> NCHW fp32 class probabilities (Dropout2d multiplier folded in); backward writes NHWC logit gradients."""

    @staticmethod
    def forward(ctx, logits, c, drop):
        prob = ops.softmax_nhwc_fwd(logits.detach(), c, drop)
        ctx.save_for_backward(prob, drop)
        ctx.ld, ctx.dtype = logits.shape[-1], logits.dtype
        return prob

    @staticmethod
    def backward(ctx, dprob):
        prob, drop = ctx.saved_tensors
        return ops.softmax_nhwc_bwd(prob, dprob, drop, ctx.ld, ctx.dtype), None, None


class SoftMaxConv(_E._B200Module):
    """train/erfnet_blocks.py:92-110: ConvTranspose2d(in, classes, 2, stride 2) -> Dropout2d (if use_dropout) -> softmax(dim=1).
    Output: NCHW fp32 probabilities.  The transposed conv runs on the tensor cores (row-parity form), dropout and softmax are one kernel."""

    def __init__(self, in_channels, softmax_classes, late_dropout_prob, use_dropout):
        super().__init__()
        self.use_dropout = use_dropout
        self.convolution = nn.ConvTranspose2d(in_channels, softmax_classes, 2, stride=2, padding=0, output_padding=0, bias=True)
        self.dropout = nn.Dropout2d(p=late_dropout_prob)
        self.in_channels, self.softmax_classes = in_channels, softmax_classes

    def _prob(self, x, ctx):
        conv = _E._PlainConv(self.convolution, ConvSpec(self.in_channels, self.softmax_classes, 2, 2, 2, transposed=True))
        logits = _E._run(conv, x, ctx)
        drop = None
        if self.use_dropout and self.training and self.dropout.p != 0:
            fixed = _E._FIXED_DROPOUT
            if fixed is not None and self in fixed:
                drop = fixed[self].to(x.device, torch.float32).reshape(x.shape[0], self.softmax_classes).contiguous()
            else:
                ones = torch.ones(x.shape[0], self.softmax_classes, 1, 1, device=x.device)
                drop = F.dropout2d(ones, self.dropout.p, True).reshape(x.shape[0], self.softmax_classes).contiguous()
        if ctx.training:
            return _SoftmaxFn.apply(logits, self.softmax_classes, drop)
        with torch.no_grad():
            return ops.softmax_nhwc_fwd(logits, self.softmax_classes, None)

    def forward(self, input):
        ctx = _E._Ctx(self._precision(), self.training)
        return self._prob(_E._ToNHWC.apply(input, input.shape[1], ctx.dtype), ctx)
