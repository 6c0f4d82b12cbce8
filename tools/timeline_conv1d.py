"""Where a conv1d_tc launch spends its time: per-CTA SM-clock stamps (erf_conv1d_tc_set_debug) of an isolated launch, averaged over CTAs.
stamps: 0 entry, 1 after pdl_wait, 2 weights arrived (MMA warp), 3 first unit's MMAs issued, 4 last unit's MMAs issued, 5/6 epilogue group 0/1 done,
7 teardown barrier passed."""
import ctypes as C
import os
import sys

import torch

os.environ["ERF_TIMELINE"] = "1"   # the instrumented build of the library (erfnet_pytorch_b200/_lib.py)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
L.erf_conv1d_tc_set_debug.argtypes = [C.c_void_p, C.c_int]
L.erf_conv1d_tc_set_debug.restype = C.c_int
buf = torch.zeros(300 * 12, device="cuda", dtype=torch.int64)
for c, h, w in [(64, 128, 256), (128, 64, 128), (16, 256, 512)]:
    for horizontal in (False, True):
        for aux in ("none", "res", "mask", "stats"):
            x = torch.randn(6, h, w, c, device="cuda").bfloat16()
            other = torch.randn_like(x)
            wt = torch.randn(c, c, *((1, 3) if horizontal else (3, 1)), device="cuda") * 0.05
            img = ops.pack_conv1d_tc(wt)
            bias = torch.zeros(c, device="cuda")
            stats = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
            kw = dict(res=other if aux == "res" else None, mask=other if aux == "mask" else None, stats=stats if aux == "stats" else None)
            for _ in range(3):
                ops.conv1d_tc(x, img, bias, horizontal, 1, relu=True, **kw)
            torch.cuda.synchronize()
            buf.zero_()
            L.erf_conv1d_tc_set_debug(C.c_void_p(buf.data_ptr()), 1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv1d_tc(x, img, bias, horizontal, 1, relu=True, **kw)
            e1.record()
            torch.cuda.synchronize()
            L.erf_conv1d_tc_set_debug(None, 1)
            t = buf.view(300, 12)[:299, :8].cpu().double()
            t = t[t[:, 0] > 0]
            ghz = 1.92
            d = lambda a, b: ((t[:, b] - t[:, a]) / ghz / 1e3)       # us
            end_epi = torch.maximum(t[:, 5], t[:, 6])
            print(f"C={c:3d} {'H' if horizontal else 'V'} {aux:5s}: launch {e0.elapsed_time(e1) * 1e3:6.1f} us | CTAs {len(t)} | setup->wait {d(0, 1).mean():5.2f} | wait->weights {d(1, 2).mean():5.2f} "
                  f"| weights->first MMAs {d(2, 3).mean():5.2f} | first->last MMAs {d(3, 4).mean():5.2f} (max {d(3, 4).max():5.2f}) "
                  f"| last MMAs->epilogue end {((end_epi - t[:, 4]) / ghz / 1e3).mean():5.2f} | ->teardown {((t[:, 7] - end_epi) / ghz / 1e3).mean():5.2f} "
                  f"| CTA total {d(0, 7).mean():5.2f} (max {d(0, 7).max():5.2f})")
