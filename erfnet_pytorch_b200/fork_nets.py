"""Shared construction of the fork's six ``train/net_*.py`` graphs on the B200 kernels.

The six reference files differ only in
  * the encoder depth            -- 5 x nb(64) + 2 x {2,4,8,16} x nb(128)  vs  the "small" 2 + 1 x {2,4,8,16},
  * one feature or three         -- plain encoder (``layers``) vs ladder encoder (``block1`` / ``block2``, returns 1/2, 1/4, 1/8 features),
  * the decoder block            -- DecoderBlock vs DecoderBlockSmall, with or without LadderBlocks,
  * the head                     -- ConvTranspose2d(16, 1 | 2) / SoftMaxConv (regression nets) or ConvTranspose2d(16, C) logits (classification nets),
  * the class_power parameters.
Every module here keeps the reference's attribute names (= checkpoint keys) and call signatures; the per-file modules
(net_regression.py ... net_classification_ladder.py) only bind the choices.  BatchNorm is off in all fork blocks; the NHWC activations stay
on the device between blocks and one autograd node per block drives the C-ABI kernels."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import erfnet as _E
from .erfnet_blocks import DecoderBlock, DecoderBlockSmall, DownsamplerBlock, LadderBlock, SoftMaxConv, non_bottleneck_1d
from .ops import ConvSpec

DILATIONS = (2, 4, 8, 16)


def _stage64(n64, use_dropout):
    return [DownsamplerBlock(16, 64)] + [non_bottleneck_1d(64, 0.03, 1, use_dropout=use_dropout) for _ in range(n64)]


def _stage128(rounds, use_dropout):
    return [DownsamplerBlock(64, 128)] + [non_bottleneck_1d(128, 0.3, d, use_dropout=use_dropout) for _ in range(rounds) for d in DILATIONS]


class SeqEncoder(_E._B200Module):
    """train/net_regression.py:16-46 (and _small, net_classification): initial_block + ``layers``; returns the 1/8 feature."""
    N64, ROUNDS = 5, 2

    def __init__(self, use_dropout=False):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.layers = nn.ModuleList(_stage64(self.N64, use_dropout) + _stage128(self.ROUNDS, use_dropout))
        self.output_conv = nn.Conv2d(128, 1, 1, stride=1, padding=0, bias=True)   # only in encoder mode (unused by forward)
        self._in_ld = 4

    def _nhwc(self, x, ctx):
        if x.shape[1] % 8 or x.shape[2] % 8:
            raise RuntimeError(f"Encoder: H and W must be multiples of 8, got {x.shape[1]}x{x.shape[2]}")
        y = _E._run(self.initial_block, x, ctx)
        for layer in self.layers:
            y = _E._run(layer, y, ctx)
        return y

    def forward(self, input):
        ctx = _E._Ctx(self._precision(), self.training)
        return _E._ToNCHW.apply(self._nhwc(_E._ToNHWC.apply(input, 4, ctx.dtype), ctx), 128)


class LadderEncoder(_E._B200Module):
    """train/net_regression_ladder.py:16-52: ``block1`` / ``block2``; returns the 1/2, 1/4 and 1/8 resolution features."""
    N64, ROUNDS = 5, 2

    def __init__(self, use_dropout=False):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.block1 = nn.ModuleList(_stage64(self.N64, use_dropout))
        self.block2 = nn.ModuleList(_stage128(self.ROUNDS, use_dropout))
        self.output_conv = nn.Conv2d(128, 1, 1, stride=1, padding=0, bias=True)
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


class _HeadMixin:
    """The last layer of every fork decoder: SoftMaxConv (class probabilities) or a plain 2x2 stride-2 transposed conv."""

    def _make_head(self, out_channels, softmax_head, late_dropout_prob=0.0, use_dropout=False):
        self.out_channels, self._softmax_head = out_channels, softmax_head
        if softmax_head:
            self.scalar_output_conv = SoftMaxConv(16, out_channels, late_dropout_prob, use_dropout=use_dropout)
        else:
            if out_channels <= 0:      # the reference fails inside ConvTranspose2d for 0 output channels (train/main.py:835-836 asserts it)
                raise RuntimeError("the classification nets need softmax_classes > 0")
            self.scalar_output_conv = nn.ConvTranspose2d(16, out_channels, 2, stride=2, padding=0, output_padding=0, bias=True)

    def _head(self, y, ctx):
        if self._softmax_head:
            return self.scalar_output_conv._prob(y, ctx)          # NCHW fp32 probabilities
        y = _E._run(_E._PlainConv(self.scalar_output_conv, ConvSpec(16, self.out_channels, 2, 2, 2, transposed=True)), y, ctx)
        return _E._ToNCHW.apply(y, self.out_channels)


class SeqDecoder(_E._B200Module, _HeadMixin):
    """train/net_regression.py:48-72: two decoder blocks and the head."""
    BLOCK = DecoderBlock

    def _build(self, out_channels, softmax_head, late_dropout_prob, use_dropout):
        self.scalar_decoder_1 = self.BLOCK(128, 64, use_dropout=use_dropout)
        self.scalar_decoder_2 = self.BLOCK(64, 16, use_dropout=use_dropout)
        self._make_head(out_channels, softmax_head, late_dropout_prob, use_dropout)

    def _nhwc(self, x, ctx):
        y = self.scalar_decoder_1._nhwc(x, ctx)
        y = self.scalar_decoder_2._nhwc(y, ctx)
        return self._head(y, ctx)

    def forward(self, input):
        ctx = _E._Ctx(self._precision(), self.training)
        return self._nhwc(_E._ToNHWC.apply(input, input.shape[1], ctx.dtype), ctx)


class LadderDecoder(_E._B200Module, _HeadMixin):
    """train/net_regression_ladder.py:54-80: decoder blocks interleaved with LadderBlocks fed by the 1/4 and 1/2 features."""
    BLOCK = DecoderBlock

    def _build(self, out_channels, softmax_head, late_dropout_prob, use_dropout):
        self.scalar_decoder_1 = self.BLOCK(128, 64, use_dropout=use_dropout)
        self.ladder_block_1 = LadderBlock(64, use_dropout=use_dropout)
        self.scalar_decoder_2 = self.BLOCK(64, 16, use_dropout=use_dropout)
        self.ladder_block_2 = LadderBlock(16, use_dropout=use_dropout)
        self._make_head(out_channels, softmax_head, late_dropout_prob, use_dropout)

    def _nhwc3(self, e1, e2, e3, ctx):
        y = self.scalar_decoder_1._nhwc(e3, ctx)
        y = self.ladder_block_1._nhwc2(y, e2, ctx)
        y = self.scalar_decoder_2._nhwc(y, ctx)
        y = self.ladder_block_2._nhwc2(y, e1, ctx)
        return self._head(y, ctx)

    def forward(self, enc1, enc2, enc3):
        ctx = _E._Ctx(self._precision(), self.training)
        e = [_E._ToNHWC.apply(t, t.shape[1], ctx.dtype) for t in (enc1, enc2, enc3)]
        return self._nhwc3(e[0], e[1], e[2], ctx)


def class_power_init(softmax_classes, spread, lo):
    """net_regression*.py:90-99 (lo = 0.0) / net_classification*.py:82-89 (lo = 0.7): ones, or a linear spread up to 2.0."""
    if spread:
        return torch.from_numpy(np.linspace(lo, 2.0, softmax_classes, dtype="float32")).view(1, softmax_classes, 1, 1)
    return torch.ones(1, softmax_classes, 1, 1)


class ForkNet(_E._B200Module):
    """Encoder -> decoder in one NHWC pass (one layout conversion in, one out)."""
    LADDER = False

    def _features(self, input, ctx):
        x = _E.input_to_nhwc(input, 4, ctx.dtype)
        if self.LADDER:
            return self.decoder._nhwc3(*self.encoder._nhwc3(x, ctx), ctx)
        return self.decoder._nhwc(self.encoder._nhwc(x, ctx), ctx)


class RegressionNet(ForkNet):
    """``Net(encoder=None, softmax_classes=0, likelihood_loss=False, spread_class_power=False, late_dropout_prob=...)`` of the four
    net_regression*.py files: [B, 1 or 2, H, W], or (prob [B,C,H,W], class_power [1,C,1,1][, class_power_var]) for softmax_classes = C > 0."""
    ENCODER, DECODER, LATE_DROPOUT = SeqEncoder, SeqDecoder, 0.1

    def __init__(self, encoder=None, softmax_classes=0, likelihood_loss=False, spread_class_power=False, late_dropout_prob=None):
        super().__init__()
        late_dropout_prob = self.LATE_DROPOUT if late_dropout_prob is None else late_dropout_prob
        self.softmax_classes = softmax_classes
        self.likelihood_loss = likelihood_loss
        use_dropout = softmax_classes > 0        # net_regression.py:87-89
        if softmax_classes > 0:
            self.class_power = nn.Parameter(class_power_init(softmax_classes, spread_class_power, 0.0))
            self.class_power_var = nn.Parameter(torch.ones(1, softmax_classes, 1, 1))
        self.encoder = self.ENCODER(use_dropout=use_dropout) if encoder is None else encoder
        self.decoder = self.DECODER(softmax_classes, late_dropout_prob, likelihood_loss, use_dropout=use_dropout)

    def forward(self, input):
        out = self._features(input, _E._Ctx(self._precision(), self.training))
        if self.softmax_classes > 0:
            if self.likelihood_loss:
                return out, self.class_power, self.class_power_var
            return out, self.class_power
        return out


class ClassificationNet(ForkNet):
    """``Net(encoder=None, softmax_classes=0, spread_class_power=False, fix_class_power=False, late_dropout_prob=0.3)`` of
    net_classification.py / net_classification_ladder.py: (logits [B,C,H,W], class_power); no softmax in the net, no dropout."""
    ENCODER, DECODER = SeqEncoder, SeqDecoder

    def __init__(self, encoder=None, softmax_classes=0, spread_class_power=False, fix_class_power=False, late_dropout_prob=0.3):
        super().__init__()
        self.softmax_classes = softmax_classes
        if softmax_classes > 0:
            init = class_power_init(softmax_classes, spread_class_power, 0.7)
            if fix_class_power:      # net_classification.py:90-92: a constant on the GPU, absent from parameters() and state_dict()
                self.class_power = init.cuda() if torch.cuda.is_available() else init
            else:
                self.class_power = nn.Parameter(init)
        self.encoder = self.ENCODER() if encoder is None else encoder
        self.decoder = self.DECODER(softmax_classes, late_dropout_prob)

    def forward(self, input):
        out = self._features(input, _E._Ctx(self._precision(), self.training))
        if self.softmax_classes > 0:
            return out, self.class_power
        return out
