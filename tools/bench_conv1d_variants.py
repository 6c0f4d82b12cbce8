"""Cost of the fused epilogue variants of the tcgen05 conv1d kernel (plain / +residual / +mask / +stats / +colsum), batch 6."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import ops  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 6
for c, h, w in [(64, 128, 256), (128, 64, 128)]:
    for horizontal in (False, True):
        x = torch.randn(batch, h, w, c, device="cuda").bfloat16()
        aux = torch.randn_like(x)
        wt = torch.randn(c, c, *((1, 3) if horizontal else (3, 1)), device="cuda") * 0.05
        img = ops.pack_conv1d_tc(wt)
        bias = torch.zeros(c, device="cuda")
        stats = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
        colsum = torch.zeros(c, device="cuda")
        variants = {"plain": {}, "res": dict(res=aux), "mask": dict(mask=aux), "stats": dict(stats=stats), "colsum": dict(colsum=colsum),
                    "mask+colsum": dict(mask=aux, colsum=colsum)}
        line = f"B={batch} C={c:3d} {'H' if horizontal else 'V'}:"
        for name, kw in variants.items():
            for _ in range(3):
                ops.conv1d_tc(x, img, bias, horizontal, 1, relu=True, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 20
            torch.cuda._sleep(int(2e7))
            e0.record()
            for _ in range(n):
                ops.conv1d_tc(x, img, bias, horizontal, 1, relu=True, **kw)
            e1.record()
            torch.cuda.synchronize()
            line += f"  {name} {e0.elapsed_time(e1) * 1e3 / n:6.1f} us"
        print(line)
