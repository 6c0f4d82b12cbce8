"""The fork's ``train/net_regression_ladder_small.py`` on the B200 kernels: same class names (Encoder, Decoder, Net), constructor signatures, attribute names
(= checkpoint keys) and return values; the small three-scale encoder and DecoderBlockSmall + LadderBlock decoder.  The shared construction lives in fork_nets.py."""
from __future__ import annotations

from . import fork_nets as _F
from .erfnet_blocks import DecoderBlockSmall


class Encoder(_F.LadderEncoder):
    """train/net_regression_ladder_small.py:16-52."""
    N64, ROUNDS = 2, 1


class Decoder(_F.LadderDecoder):
    """train/net_regression_ladder_small.py:54-80."""
    BLOCK = DecoderBlockSmall

    def __init__(self, softmax_classes, late_dropout_prob, likelihood_loss, use_dropout=False):
        super().__init__()
        self.softmax_classes = softmax_classes
        if softmax_classes:
            self._build(softmax_classes, True, late_dropout_prob, use_dropout)
        else:
            self._build(2 if likelihood_loss else 1, False, late_dropout_prob, use_dropout)   # mean (and std) of the regressed quantity


class Net(_F.RegressionNet):
    """train/net_regression_ladder_small.py:83-127."""
    ENCODER, DECODER, LADDER = Encoder, Decoder, True
    LATE_DROPOUT = 0.0
