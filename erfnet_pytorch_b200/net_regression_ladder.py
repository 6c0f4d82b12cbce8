"""The fork's ladder regression net (train/net_regression_ladder.py, BASELINE config 4) on the B200 kernels: same class names,
constructor signatures, attribute names (= checkpoint keys: encoder.initial_block / block1 / block2 / output_conv,
decoder.scalar_decoder_1/2, ladder_block_1/2, scalar_output_conv) and return values.  BatchNorm is off in all blocks and dropout is
on only when ``softmax_classes > 0`` (net_regression_ladder.py:95-97), so the training graph is conv + bias + ReLU + residual only.

With ``softmax_classes = C > 0`` the head is a SoftMaxConv (class probabilities) and the net owns the ``class_power`` /
``class_power_var`` parameters consumed by the class-probability losses (train/main.py:327-338)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import erfnet as _E
from .erfnet_blocks import DecoderBlock, DownsamplerBlock, LadderBlock, SoftMaxConv, non_bottleneck_1d
from .ops import ConvSpec


class Encoder(_E._B200Module):
    """train/net_regression_ladder.py:16-52: returns the 1/2, 1/4 and 1/8 resolution features."""

    def __init__(self, use_dropout=False):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.block1 = nn.ModuleList()
        self.block2 = nn.ModuleList()
        self.block1.append(DownsamplerBlock(16, 64))
        for _ in range(0, 5):
            self.block1.append(non_bottleneck_1d(64, 0.03, 1, use_dropout=use_dropout))
        self.block2.append(DownsamplerBlock(64, 128))
        for _ in range(0, 2):
            self.block2.append(non_bottleneck_1d(128, 0.3, 2, use_dropout=use_dropout))
            self.block2.append(non_bottleneck_1d(128, 0.3, 4, use_dropout=use_dropout))
            self.block2.append(non_bottleneck_1d(128, 0.3, 8, use_dropout=use_dropout))
            self.block2.append(non_bottleneck_1d(128, 0.3, 16, use_dropout=use_dropout))
        self.output_conv = nn.Conv2d(128, 1, 1, stride=1, padding=0, bias=True)   # only in encoder mode (unused by forward)
        self._in_ld = 4

    def _nhwc3(self, x, ctx):
        if x.shape[1] % 8 or x.shape[2] % 8:
            raise RuntimeError(f"Encoder: H and W must be multiples of 8, got {x.shape[1]}x{x.shape[2]}")
        o1 = _E._run(self.initial_block, x, ctx)
        o2 = o1
        for layer in self.block1:
            o2 = _E._run(layer, o2, ctx)
        o3 = o2
        for layer in self.block2:
            o3 = _E._run(layer, o3, ctx)
        return o1, o2, o3

    def forward(self, input):
        ctx = _E._Ctx(self._precision(), self.training)
        o1, o2, o3 = self._nhwc3(_E._ToNHWC.apply(input, 4, ctx.dtype), ctx)
        return _E._ToNCHW.apply(o1, 16), _E._ToNCHW.apply(o2, 64), _E._ToNCHW.apply(o3, 128)


class Decoder(_E._B200Module):
    """train/net_regression_ladder.py:54-80."""

    def __init__(self, softmax_classes, late_dropout_prob, likelihood_loss, use_dropout=False):
        super().__init__()
        self.softmax_classes = softmax_classes
        self.scalar_decoder_1 = DecoderBlock(128, 64, use_dropout=use_dropout)
        self.ladder_block_1 = LadderBlock(64, use_dropout=use_dropout)
        self.scalar_decoder_2 = DecoderBlock(64, 16, use_dropout=use_dropout)
        self.ladder_block_2 = LadderBlock(16, use_dropout=use_dropout)
        if softmax_classes:
            self.out_channels = softmax_classes
            self.scalar_output_conv = SoftMaxConv(16, softmax_classes, late_dropout_prob, use_dropout=use_dropout)
        else:
            self.out_channels = 2 if likelihood_loss else 1       # mean (and std) of the regressed quantity
            self.scalar_output_conv = nn.ConvTranspose2d(16, self.out_channels, 2, stride=2, padding=0, output_padding=0, bias=True)

    def _nhwc3(self, e1, e2, e3, ctx):
        y = self.scalar_decoder_1._nhwc(e3, ctx)
        y = self.ladder_block_1._nhwc2(y, e2, ctx)
        y = self.scalar_decoder_2._nhwc(y, ctx)
        y = self.ladder_block_2._nhwc2(y, e1, ctx)
        if self.softmax_classes:
            return self.scalar_output_conv._prob(y, ctx)          # NCHW fp32 probabilities
        y = _E._run(_E._PlainConv(self.scalar_output_conv, ConvSpec(16, self.out_channels, 2, 2, 2, transposed=True)), y, ctx)
        return _E._ToNCHW.apply(y, self.out_channels)

    def forward(self, enc1, enc2, enc3):
        ctx = _E._Ctx(self._precision(), self.training)
        e = [_E._ToNHWC.apply(t, t.shape[1], ctx.dtype) for t in (enc1, enc2, enc3)]
        return self._nhwc3(e[0], e[1], e[2], ctx)


class Net(_E._B200Module):
    """train/net_regression_ladder.py:83-128: ``Net(...)(input)`` -> [B, 1 or 2, H, W], or with softmax_classes = C > 0
    (prob [B,C,H,W], class_power [1,C,1,1][, class_power_var])."""

    def __init__(self, encoder=None, softmax_classes=0, likelihood_loss=False, spread_class_power=False, late_dropout_prob=0.0):
        super().__init__()
        self.softmax_classes = softmax_classes
        self.likelihood_loss = likelihood_loss
        use_dropout = softmax_classes > 0
        if softmax_classes > 0:   # train/net_regression_ladder.py:98-111
            if spread_class_power:
                init = torch.from_numpy(np.linspace(0.0, 2.0, softmax_classes, dtype="float32")).view(1, softmax_classes, 1, 1)
            else:
                init = torch.ones(1, softmax_classes, 1, 1)
            self.class_power = nn.Parameter(init)
            self.class_power_var = nn.Parameter(torch.ones(1, softmax_classes, 1, 1))
        self.encoder = Encoder(use_dropout=use_dropout) if encoder is None else encoder
        self.decoder = Decoder(softmax_classes, late_dropout_prob, likelihood_loss, use_dropout=use_dropout)

    def forward(self, input):
        ctx = _E._Ctx(self._precision(), self.training)
        x = _E._ToNHWC.apply(input, 4, ctx.dtype)
        e1, e2, e3 = self.encoder._nhwc3(x, ctx)
        out = self.decoder._nhwc3(e1, e2, e3, ctx)
        if self.softmax_classes > 0:
            if self.likelihood_loss:
                return out, self.class_power, self.class_power_var
            return out, self.class_power
        return out
