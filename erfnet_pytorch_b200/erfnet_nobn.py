"""eval/erfnet_nobn.py on the B200 kernels: the ERFNet graph "without batch normalization layers nor dropout" that eval_forwardTime.py
times (l.15, 23).  In the reference the BatchNorm / Dropout2d modules are still CONSTRUCTED (so the checkpoint keys are those of
eval/erfnet.py) but never called (erfnet_nobn.py:23, 51, 57, 60, 109); the encoder's drop probabilities are 0.03 / 0.3 (l.76-84).  Here that
is the Cityscapes graph assembled from the BN-free, dropout-free blocks of erfnet_blocks.py."""
from __future__ import annotations

import torch.nn as nn

from . import erfnet as _E
from .erfnet_blocks import DownsamplerBlock, UpsamplerBlock, non_bottleneck_1d
from .ops import ConvSpec


class Encoder(_E.Encoder):
    """eval/erfnet_nobn.py:65-97."""

    def __init__(self, num_classes):
        _E._B200Module.__init__(self)
        self.initial_block = DownsamplerBlock(3, 16)
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(16, 64))
        for _ in range(0, 5):
            self.layers.append(non_bottleneck_1d(64, 0.03, 1))
        self.layers.append(DownsamplerBlock(64, 128))
        for _ in range(0, 2):
            for d in (2, 4, 8, 16):
                self.layers.append(non_bottleneck_1d(128, 0.3, d))
        self.output_conv = nn.Conv2d(128, num_classes, 1, stride=1, padding=0, bias=True)
        self.num_classes = num_classes
        self._in_ld = 4


class Decoder(_E.Decoder):
    """eval/erfnet_nobn.py:112-136."""

    def __init__(self, num_classes):
        _E._B200Module.__init__(self)
        self.layers = nn.ModuleList()
        self.layers.append(UpsamplerBlock(128, 64))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(UpsamplerBlock(64, 16))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.output_conv = nn.ConvTranspose2d(16, num_classes, 2, stride=2, padding=0, output_padding=0, bias=True)
        self.num_classes = num_classes
        self._out_c = num_classes


class ERFNet(_E.ERFNet):
    """eval/erfnet_nobn.py:139-154."""

    def __init__(self, num_classes, encoder=None):
        _E._B200Module.__init__(self)
        self.encoder = Encoder(num_classes) if encoder is None else encoder
        self.decoder = Decoder(num_classes)
        self.num_classes = num_classes
