"""Where a conv1d_pair_tc launch (eval mode: 3x1 -> ReLU -> 1x3 in one kernel) spends its time: per-CTA SM-clock stamps of an isolated launch with
cold inputs (tensor sets rotating over > L2), averaged over CTAs.  stamps: 0 entry, 1 after griddepcontrol.wait, 2 weights arrived, 3 first row's
stage-1 MMAs issued, 4 last MMAs issued, 5 stage-1 epilogue done, 6 stage-2 epilogue done, 7 exit."""
import ctypes as C
import os
import sys

import torch

os.environ["ERF_TIMELINE"] = "1"   # the instrumented build of the library (erfnet_pytorch_b200/_lib.py)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
L.erf_conv1d_pair_tc_set_debug.argtypes = [C.c_void_p]
L.erf_conv1d_pair_tc_set_debug.restype = C.c_int
buf = torch.zeros(296 * 8, device="cuda", dtype=torch.int64)
ghz = 1.965
for batch in (6, 32):
    for c, h, w in [(64, 128, 256), (16, 256, 512)]:
        for with_res in (False, True):
            nsets = max(2, int(300e6 // (batch * h * w * c * 2 * 2)))
            xs = [torch.randn(batch, h, w, c, device="cuda").bfloat16() for _ in range(nsets)]
            w1 = ops.pack_conv1d_tc(torch.randn(c, c, 3, 1, device="cuda") * 0.05)
            w2 = ops.pack_conv1d_tc(torch.randn(c, c, 1, 3, device="cuda") * 0.05)
            b1, b2 = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
            run = lambda i: ops.conv1d_pair_tc(xs[i % nsets], w1, b1, w2, b2, 1, relu=True, res=xs[(i + 1) % nsets] if with_res else None, static_w=True)
            for i in range(4):
                run(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(int(2e7))
            e0.record()
            for i in range(16):
                run(i)
            e1.record()
            torch.cuda.synchronize()
            us16 = e0.elapsed_time(e1) * 1e3 / 16
            buf.zero_()
            L.erf_conv1d_pair_tc_set_debug(C.c_void_p(buf.data_ptr()))
            run(5)
            torch.cuda.synchronize()
            L.erf_conv1d_pair_tc_set_debug(None)
            t = buf.view(296, 8).cpu().double()
            t = t[t[:, 0] > 0]
            d = lambda a, b: float(((t[:, b] - t[:, a]) / ghz / 1e3).mean())
            rows = batch * h / len(t)
            print(f"B={batch:2d} C={c:2d} {h}x{w} res={int(with_res)}: {us16:6.1f} us/launch (16 back to back, cold) | CTAs {len(t)} rows/CTA {rows:5.1f} | setup->wait {d(0, 1):5.2f} "
                  f"| +weights {d(1, 2):5.2f} | +first row S1 {d(2, 3):5.2f} | MMA loop {d(3, 4):6.2f} ({d(3, 4) / max(rows - 1, 1):5.2f}/row) | last MMA->E1 end {d(4, 5):5.2f} "
                  f"->E2 end {d(4, 6):5.2f} | CTA total {d(0, 7):6.2f}")
