"""Where a fused BatchNorm-backward launch spends its time: per-CTA globaltimer stamps (erf_bn_set_debug) of one cold launch.
stamps: 0 entry, 1 after griddepcontrol.wait, 2 phase-1 loads consumed, 3 block reduce + fp64 atomics issued, 4 grid barrier passed,
5 coefficients ready, 6 phase 2 done.  Times in us relative to the earliest CTA entry; min / mean / max over CTAs."""
import ctypes as C
import os
import sys

import torch

os.environ["ERF_TIMELINE"] = "1"   # the instrumented build of the library (erfnet_pytorch_b200/_lib.py)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import _lib, ops  # noqa: E402

L = _lib.lib()
L.erf_bn_set_debug.argtypes = [C.c_void_p]
L.erf_bn_set_debug.restype = C.c_int
buf = torch.zeros(296 * 8, device="cuda", dtype=torch.int64)
arena = ops.ScratchArena(torch.device("cuda"), n=1 << 20)
ops.ARENA = arena
for c, h, w in [(64, 128, 256), (128, 64, 128), (16, 256, 512)]:
    b = 6
    sets = [[torch.randn(b, h, w, c, device="cuda").bfloat16() for _ in range(3)] for _ in range(6)]
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.rand(c, device="cuda")
    mean, invstd = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
    drop = (torch.rand(b, c, device="cuda") > 0.1).float() / 0.9
    for mode in ("y-mask", "tensor-mask", "drop+gm"):
        def run(i):
            g, y, m = sets[i % 6]
            ops.bn_bwd(g, y, gamma, mean, invstd, mask=m, drop=drop if mode == "drop+gm" else None, want_masked=mode == "drop+gm",
                       beta_for_mask=beta if mode == "y-mask" else None)
        arena.reset()
        for i in range(5):
            run(i)
        torch.cuda.synchronize()
        buf.zero_()
        L.erf_bn_set_debug(C.c_void_p(buf.data_ptr()))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(5)
        e1.record()
        torch.cuda.synchronize()
        L.erf_bn_set_debug(None)
        t = buf.view(296, 8).cpu().double()
        t = t[t[:, 0] > 0]
        t0 = t[:, 0].min()
        rel = (t[:, :7] - t0) / 1e3
        names = ["entry", "pdl-wait", "phase1", "reduce+atomics", "barrier", "coefs", "phase2"]
        line = " | ".join(f"{n} {rel[:, k].min():5.1f}/{rel[:, k].mean():5.1f}/{rel[:, k].max():5.1f}" for k, n in enumerate(names))
        print(f"C={c:3d} {mode:11s}: launch {e0.elapsed_time(e1) * 1e3:6.1f} us (event pair) | CTAs {len(t)} | {line}")
ops.ARENA = None
