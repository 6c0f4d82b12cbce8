"""Isolated timing of the BatchNorm backward kernels (reduce + apply) and affine_act at the bench shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda._sleep(int(2e7))
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n


for c, h, w in [(64, 128, 256), (128, 64, 128), (16, 256, 512)]:
    b = 6
    # rotate over 6 tensor sets so that the working set (6 x 4 x T) exceeds L2 like in the real step
    sets = [[torch.randn(b, h, w, c, device="cuda").bfloat16() for _ in range(3)] for _ in range(6)]
    gamma, mean, invstd = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
    drop = (torch.rand(b, c, device="cuda") > 0.1).float() / 0.9
    it = [0]

    def run(masked):
        g, y, m = sets[it[0] % 6]
        it[0] += 1
        ops.bn_bwd(g, y, gamma, mean, invstd, mask=m, drop=drop, want_masked=masked)

    t_bytes = b * h * w * c * 2
    from erfnet_pytorch_b200 import _lib
    L = _lib.lib()
    sums = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
    dy = torch.empty_like(sets[0][0])
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    st = ops._stream(dy)

    def red():
        g, y, m = sets[it[0] % 6]
        it[0] += 1
        ops.check(L.erf_bn_bwd_reduce(ops._p(g), c, ops._p(y), c, ops._p(m), c, ops._p(drop), ops._p(mean), ops._p(invstd), b, h * w, c, ops._p(sums), 1, st))

    def app():
        g, y, m = sets[it[0] % 6]
        it[0] += 1
        ops.check(L.erf_bn_bwd_apply(ops._p(g), c, ops._p(y), c, ops._p(m), c, ops._p(drop), ops._p(gamma), ops._p(mean), ops._p(invstd), ops._p(sums), b, h * w, c,
                                     ops._p(dy), c, ops._p(dg), ops._p(db), None, c, 1, st))

    ur, ua = timeit(red), timeit(app)
    print(f"C={c:3d}: reduce {ur:6.1f} us ({3 * t_bytes / ur / 1e6:5.2f} TB/s)   apply {ua:6.1f} us ({4 * t_bytes / ua / 1e6:5.2f} TB/s)")
    us = timeit(lambda: run(False))
    us2 = timeit(lambda: run(True))
    sc, sh = torch.rand(c, device="cuda"), torch.rand(c, device="cuda")

    def aff():
        g, y, m = sets[it[0] % 6]
        it[0] += 1
        ops.affine_act(g, sc, sh, res=y, relu=True, out=m)

    us3 = timeit(aff)
    print(f"C={c:3d} {h}x{w}: bn_bwd {us:6.1f} us ({7 * t_bytes / us / 1e6:5.2f} TB/s)  +gmasked {us2:6.1f} us ({8 * t_bytes / us2 / 1e6:5.2f} TB/s)  "
          f"affine_act(res) {us3:6.1f} us ({3 * t_bytes / us3 / 1e6:5.2f} TB/s)")
