"""The fork's ``train/net_regression_small.py`` on the B200 kernels: same class names (Encoder, Decoder, Net), constructor signatures, attribute names
(= checkpoint keys) and return values; the small encoder (2 x nb(64), one {2,4,8,16} round of nb(128)) and two DecoderBlockSmall.  The shared construction lives in fork_nets.py."""
from __future__ import annotations

from . import fork_nets as _F
from .erfnet_blocks import DecoderBlockSmall


class Encoder(_F.SeqEncoder):
    """train/net_regression_small.py:16-46."""
    N64, ROUNDS = 2, 1


class Decoder(_F.SeqDecoder):
    """train/net_regression_small.py:48-72."""
    BLOCK = DecoderBlockSmall

    def __init__(self, softmax_classes, late_dropout_prob, likelihood_loss, use_dropout=False):
        super().__init__()
        self.softmax_classes = softmax_classes
        if softmax_classes:
            self._build(softmax_classes, True, late_dropout_prob, use_dropout)
        else:
            self._build(2 if likelihood_loss else 1, False, late_dropout_prob, use_dropout)   # mean (and std) of the regressed quantity


class Net(_F.RegressionNet):
    """train/net_regression_small.py:75-119."""
    ENCODER, DECODER, LADDER = Encoder, Decoder, False
    LATE_DROPOUT = 0.1
