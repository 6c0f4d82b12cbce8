import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import erfnet as E, ops
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input

def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum())

for (c, d, hw) in [(64, 1, (8, 256)), (128, 4, (12, 128))]:
    blk = E.non_bottleneck_1d(c, 0.1, d)
    bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=3)
    x = golden_input((2, c) + hw, 40 + d) - 0.3
    gy = golden_input((2, c) + hw, 41) - 0.5
    keep = (torch.rand(2, c, generator=torch.Generator().manual_seed(1)) > 0.1).float() / 0.9
    sdo = {k: v.clone() for k, v in bsd.items()}
    for k, v in sdo.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    yo = O.non_bottleneck_1d(xo, sdo, "blk", d, 0.1, training=True, drop_scale=keep.view(2, c, 1, 1))
    yo.backward(gy)
    for prec, use_tc in (("fp32", False), ("bf16", False), ("bf16", True)):
        blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
        blk.cuda().train()
        blk.zero_grad()
        ops.USE_TC = use_tc
        dtp = torch.float32 if prec == "fp32" else torch.bfloat16
        ctx = E._Ctx(prec, True, {blk: keep})
        xn = ops.nchw_to_nhwc(x.cuda(), c, dtp).requires_grad_(True)
        yn = E._run(blk, xn, ctx)
        yn.backward(ops.nchw_to_nhwc(gy.cuda(), c, dtp))
        print(c, d, prec, "tc" if use_tc else "gen", "fwd %.3e dx %.3e dW4 %.3e dW1 %.3e dg1 %.3e" % (
            rel(ops.nhwc_to_nchw(yn.detach(), c).cpu(), yo.detach()), rel(ops.nhwc_to_nchw(xn.grad, c).cpu(), xo.grad),
            rel(blk.conv1x3_2.weight.grad.cpu(), sdo["blk.conv1x3_2.weight"].grad), rel(blk.conv3x1_1.weight.grad.cpu(), sdo["blk.conv3x1_1.weight"].grad),
            rel(blk.bn1.weight.grad.cpu(), sdo["blk.bn1.weight"].grad)))
