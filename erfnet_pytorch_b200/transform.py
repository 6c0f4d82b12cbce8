"""Device-side versions of the label / image transforms either side of the model (eval/transform.py = train/transform.py,
eval/eval_iou.py:30-38, eval/eval_cityscapes_server.py:39-63): same class names and call signatures, but they take and return CUDA tensors
and each is one kernel of liberfnet_b200.so.  PIL decoding / resizing stays on the host (out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def colormap_cityscapes(n):
    """eval/transform.py:10-33: the 19 Cityscapes train-id colours, the rest black."""
    cmap = np.zeros([n, 3]).astype(np.uint8)
    table = [(128, 64, 128), (244, 35, 232), (70, 70, 70), (102, 102, 156), (190, 153, 153), (153, 153, 153), (250, 170, 30), (220, 220, 0),
             (107, 142, 35), (152, 251, 152), (70, 130, 180), (220, 20, 60), (255, 0, 0), (0, 0, 142), (0, 0, 70), (0, 60, 100), (0, 80, 100),
             (0, 0, 230), (119, 11, 32), (0, 0, 0)]
    for i, rgb in enumerate(table[:n]):
        cmap[i, :] = np.array(rgb)
    return cmap


class ToLabel:
    """eval/transform.py:67-70: image (uint8 [H,W] or [N,H,W] tensor on the device) -> int64 labels with a leading channel axis."""

    def __call__(self, image):
        t = image if torch.is_tensor(image) else torch.from_numpy(np.array(image))
        lut = torch.arange(256, device=t.device, dtype=torch.int64)
        return ops.label_lut(t, lut).unsqueeze(0 if t.dim() == 2 else 1)


class Relabel:
    """eval/transform.py:55-64: tensor[tensor == olabel] = nlabel (returns a new int64 tensor)."""

    def __init__(self, olabel, nlabel):
        self.olabel, self.nlabel = olabel, nlabel
        self._lut = {}

    def __call__(self, tensor):
        if tensor.dtype not in (torch.int64, torch.uint8):
            raise AssertionError("tensor needs to be LongTensor")
        lut = self._lut.get(tensor.device)
        if lut is None:
            lut = torch.arange(max(256, self.olabel + 1), dtype=torch.int64)
            lut[self.olabel] = self.nlabel
            lut = self._lut[tensor.device] = lut.to(tensor.device)
        return ops.label_lut(tensor, lut)


class LabelTable:
    """A general label look-up table, e.g. train id -> label id before writing server submissions (eval_cityscapes_server.py:39-63)."""

    def __init__(self, table):
        self.table = torch.as_tensor(table, dtype=torch.int64)
        self._dev = {}

    def __call__(self, tensor):
        lut = self._dev.get(tensor.device)
        if lut is None:
            lut = self._dev[tensor.device] = self.table.to(tensor.device)
        return ops.label_lut(tensor, lut)


class Colorize:
    """eval/transform.py:73-93: label image [1,H,W] (or a batch [N,H,W]) -> uint8 colour image [3,H,W] ([N,3,H,W])."""

    def __init__(self, n=22):
        cmap = colormap_cityscapes(256)
        cmap[n] = cmap[-1]
        self.cmap = torch.from_numpy(cmap[:n])
        self._dev = {}

    def __call__(self, gray_image):
        cmap = self._dev.get(gray_image.device)
        if cmap is None:
            cmap = self._dev[gray_image.device] = self.cmap.to(gray_image.device).contiguous()
        single = gray_image.shape[0] == 1 and gray_image.dim() == 3
        out = ops.colorize(gray_image if gray_image.dim() == 3 else gray_image[:, 0], cmap)
        return out[0] if single else out
