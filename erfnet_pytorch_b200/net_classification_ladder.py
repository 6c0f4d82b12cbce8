"""The fork's ``train/net_classification_ladder.py`` on the B200 kernels: same class names (Encoder, Decoder, Net), constructor signatures, attribute names
(= checkpoint keys) and return values; three-scale encoder, DecoderBlock + LadderBlock decoder and a logits head; returns (logits, class_power).  The shared construction lives in fork_nets.py."""
from __future__ import annotations

from . import fork_nets as _F
from .erfnet_blocks import DecoderBlock


class Encoder(_F.LadderEncoder):
    """train/net_classification_ladder.py:16-52."""
    N64, ROUNDS = 5, 2

    def __init__(self):
        super().__init__(use_dropout=False)


class Decoder(_F.LadderDecoder):
    """train/net_classification_ladder.py:55-74: the head is a plain ConvTranspose2d(16, softmax_classes, 2, stride 2) -> logits."""
    BLOCK = DecoderBlock

    def __init__(self, softmax_classes, late_dropout_prob):
        super().__init__()
        self.softmax_classes = softmax_classes
        self._build(softmax_classes, False, late_dropout_prob, False)


class Net(_F.ClassificationNet):
    """train/net_classification_ladder.py:77-117."""
    ENCODER, DECODER, LADDER = Encoder, Decoder, True
