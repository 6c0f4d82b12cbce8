"""The fork's ``train/net_regression_ladder.py`` (BASELINE config 4) on the B200 kernels: same class names (Encoder, Decoder, Net), constructor
signatures, attribute names (= checkpoint keys: encoder.initial_block / block1 / block2 / output_conv, decoder.scalar_decoder_1/2,
ladder_block_1/2, scalar_output_conv) and return values; three-scale encoder (5 x nb(64), 2 x {2,4,8,16} x nb(128)) and DecoderBlock + LadderBlock
decoder.  BatchNorm is off in all blocks and dropout is on only when ``softmax_classes > 0`` (net_regression_ladder.py:95-97).  With
``softmax_classes = C > 0`` the head is a SoftMaxConv and the net owns the ``class_power`` / ``class_power_var`` parameters consumed by the
class-probability losses (train/main.py:327-338).  The shared construction lives in fork_nets.py."""
from __future__ import annotations

from . import fork_nets as _F
from .erfnet_blocks import DecoderBlock


class Encoder(_F.LadderEncoder):
    """train/net_regression_ladder.py:16-52: returns the 1/2, 1/4 and 1/8 resolution features."""
    N64, ROUNDS = 5, 2


class Decoder(_F.LadderDecoder):
    """train/net_regression_ladder.py:54-80."""
    BLOCK = DecoderBlock

    def __init__(self, softmax_classes, late_dropout_prob, likelihood_loss, use_dropout=False):
        super().__init__()
        self.softmax_classes = softmax_classes
        if softmax_classes:
            self._build(softmax_classes, True, late_dropout_prob, use_dropout)
        else:
            self._build(2 if likelihood_loss else 1, False, late_dropout_prob, use_dropout)   # mean (and std) of the regressed quantity


class Net(_F.RegressionNet):
    """train/net_regression_ladder.py:83-128: ``Net(...)(input)`` -> [B, 1 or 2, H, W], or with softmax_classes = C > 0
    (prob [B,C,H,W], class_power [1,C,1,1][, class_power_var])."""
    ENCODER, DECODER, LADDER = Encoder, Decoder, True
    LATE_DROPOUT = 0.0
