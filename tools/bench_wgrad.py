"""Isolated timing of the tcgen05 3-tap weight-gradient kernel (+ its reduction, whichever form is active) per shape class:
CUDA events over 20 back-to-back launches.  ERF_WGRAD_BULK=0 selects the per-CTA partial slabs + reduce kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import ops  # noqa: E402

cases = [(64, 128, 256), (128, 64, 128), (16, 256, 512)]
for batch in (6, 24):
    for c, h, w in cases:
        for horizontal in (False, True):
            x = torch.randn(batch, h, w, c, device="cuda").bfloat16()
            dy = torch.randn_like(x)
            dw = torch.zeros(c, c, *((1, 3) if horizontal else (3, 1)), device="cuda")
            for _ in range(3):
                ops.conv1d_wgrad_tc(x, dy, horizontal, 1, dw_out=dw, want_bias=False)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            torch.cuda._sleep(int(2e7))
            e0.record()
            for _ in range(n):
                ops.conv1d_wgrad_tc(x, dy, horizontal, 1, dw_out=dw, want_bias=False)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / n
            fl = 2.0 * batch * h * w * c * c * 3
            print(f"B={batch:2d} C={c:3d} {h}x{w} {'H' if horizontal else 'V'}: {us:7.1f} us  {fl / us / 1e6:7.1f} TF/s  {2 * x.numel() * 2 / us / 1e3:7.1f} GB/s (algorithmic)")
