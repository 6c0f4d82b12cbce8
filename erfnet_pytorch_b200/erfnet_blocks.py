"""The fork's block module (train/erfnet_blocks.py) on the B200 kernels: same class names and constructor signatures,
with the FORK's defaults -- BatchNorm off, dropout off unless ``use_dropout`` -- so that ``from erfnet_blocks import *``
in the fork's net_*.py files resolves to these.  The BatchNorm modules are still constructed (their parameters and
buffers are part of the fork's checkpoints) but take no part in the arithmetic when ``use_batchnorm`` is False.

Built so far: DownsamplerBlock, non_bottleneck_1d, UpsamplerBlock, DecoderBlockSmall, DecoderBlock (train + eval, all
on the same kernels as erfnet.py).  SoftMaxConv and LadderBlock are SURVEY section 8(f) "next" items."""
from __future__ import annotations

import torch.nn as nn

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
