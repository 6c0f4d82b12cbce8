"""The fork's ``train/net_classification.py`` on the B200 kernels: same class names (Encoder, Decoder, Net), constructor signatures, attribute names
(= checkpoint keys) and return values; plain encoder, two DecoderBlocks and a logits head; returns (logits, class_power).  The shared construction lives in fork_nets.py."""
from __future__ import annotations

from . import fork_nets as _F
from .erfnet_blocks import DecoderBlock


class Encoder(_F.SeqEncoder):
    """train/net_classification.py:17-47."""
    N64, ROUNDS = 5, 2

    def __init__(self):
        super().__init__(use_dropout=False)


class Decoder(_F.SeqDecoder):
    """train/net_classification.py:49-66: the head is a plain ConvTranspose2d(16, softmax_classes, 2, stride 2) -> logits."""
    BLOCK = DecoderBlock

    def __init__(self, softmax_classes, late_dropout_prob):
        super().__init__()
        self.softmax_classes = softmax_classes
        self._build(softmax_classes, False, late_dropout_prob, False)


class Net(_F.ClassificationNet):
    """train/net_classification.py:69-110."""
    ENCODER, DECODER, LADDER = Encoder, Decoder, False
