"""Isolated timing of the train-mode BatchNorm kernels at the bench shapes: backward as one launch (csrc/bn_fused.cu) against the round-1
reduce + apply pair, forward apply as one launch against finalize + affine_act; cold (tensor sets rotating over > L2) and hot (one set)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import ops  # noqa: E402


def timeit(fn, n=24):
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


arena = ops.ScratchArena(torch.device("cuda"), n=1 << 20)
ops.ARENA = arena
for c, h, w in [(64, 128, 256), (128, 64, 128), (16, 256, 512)]:
    b = 6
    sets = [[torch.randn(b, h, w, c, device="cuda").bfloat16() for _ in range(3)] for _ in range(6)]
    gamma, beta = torch.rand(c, device="cuda") + 0.5, torch.rand(c, device="cuda")
    mean, invstd = torch.randn(c, device="cuda") * 0.1, torch.rand(c, device="cuda") + 0.5
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    drop = (torch.rand(b, c, device="cuda") > 0.1).float() / 0.9
    stats = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
    ops.channel_stats(sets[0][0], stats)
    it = [0]
    t_bytes = b * h * w * c * 2
    for hot in (False, True):
        def pick():
            s = sets[0] if hot else sets[it[0] % 6]
            it[0] += 1
            return s

        def bwd(masked, from_y=False):
            g, y, m = pick()
            if arena.cursor + 2 * c + 1 > arena.buf.numel():
                arena.cursor = 0      # (the slots are only ever incremented by the barrier: reuse is fine for timing as long as they were re-zeroed)
                arena.buf.zero_()
            ops.bn_bwd(g, y, gamma, mean, invstd, mask=m, drop=drop if masked else None, want_masked=masked, beta_for_mask=beta if from_y else None)

        def fwd(res):
            g, y, m = pick()
            ops.bn_apply(g, stats, b * h * w, gamma, beta, rm, rv, 0.1, 1e-3, drop=drop if res else None, res=y if res else None, relu=True)

        row = []
        for fused in (True, False):
            ops.BN_FUSED = fused
            arena.reset()
            row += [timeit(lambda: bwd(False)), timeit(lambda: bwd(True)), timeit(lambda: fwd(False)), timeit(lambda: fwd(True))]
        ops.BN_FUSED = True
        arena.reset()
        fy = timeit(lambda: bwd(False, True))
        f = row[:4]
        u = row[4:]
        print(f"C={c:3d} {h}x{w} {'hot ' if hot else 'cold'}: bn_bwd fused {f[0]:6.1f} us ({4 * t_bytes / f[0] / 1e6:5.2f} TB/s algorithmic), mask from y {fy:6.1f} us / two launches {u[0]:6.1f} us;  "
              f"+drop+gmasked {f[1]:6.1f} / {u[1]:6.1f} us;  bn_apply fused {f[2]:6.1f} / finalize+affine {u[2]:6.1f} us;  +drop+res {f[3]:6.1f} / {u[3]:6.1f} us")
ops.ARENA = None
