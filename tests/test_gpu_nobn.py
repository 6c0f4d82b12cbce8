"""GPU: the fork's BN-free blocks (train/erfnet_blocks.py defaults) in train mode through the C ABI, fp32 check mode
(generic kernels, tight) and bf16 (tcgen05 kernels at TC-eligible sizes) against the reference's golden outputs."""
import numpy as np
import pytest
import torch

from erfnet_pytorch_b200 import erfnet as E
from erfnet_pytorch_b200 import erfnet_blocks as EB
from erfnet_pytorch_b200 import ops
from oracle import erfnet_oracle as O
from tests.nobn_cases import assert_rel_l2, check_blocks_nobn
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu


def test_blocks_nobn_fp32(golden_dir):
    check_blocks_nobn(golden_dir, "cuda", "fp32", tol=1e-3)


def test_blocks_nobn_bf16_small(golden_dir):
    # bf16 storage: ~2^-8 per tensor and a few flipped ReLU masks -> relative L2
    check_blocks_nobn(golden_dir, "cuda", "bf16", tol=0.15, l2=True)


@pytest.mark.parametrize("c,d,p", [(64, 1, 0.3), (128, 8, 0.1), (16, 1, 0.0), (128, 2, 0.0)])
def test_nb1d_nobn_tc_vs_oracle(c, d, p):
    """TC-eligible size (W % 128 == 0): the tcgen05 path of the BN-free block vs the fp32 oracle on the same bf16-rounded
    operands, and vs the generic kernels on identical bf16 data."""
    torch.manual_seed(3)
    n, h, w = 2, 16, 128
    blk = EB.non_bottleneck_1d(c, p, d, use_dropout=True)
    assert blk.use_batchnorm is False
    sd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=21)
    blk.load_state_dict({k[4:]: v for k, v in sd.items()})
    blk.cuda().train()
    x0 = (torch.rand(n, c, h, w) - 0.3).bfloat16().float()
    gy = (torch.rand(n, c, h, w) - 0.5).bfloat16().float()
    drop = torch.nn.functional.dropout2d(torch.ones(n, c, 1, 1), p, True).reshape(n, c) if p else None
    res = {}
    for mode in ("tc", "generic"):
        ops.USE_TC = mode == "tc"
        try:
            blk.zero_grad()
            x = x0.clone().cuda().requires_grad_(True)
            with E.fixed_dropout({blk: drop} if drop is not None else {}):
                y = blk(x)
            y.backward(gy.cuda())
            res[mode] = [y.detach().float().cpu(), x.grad.cpu()] + [getattr(blk, cn).weight.grad.cpu().clone() for cn in ("conv3x1_1", "conv1x3_2")] + \
                        [getattr(blk, cn).bias.grad.cpu().clone() for cn in ("conv3x1_1", "conv1x3_1", "conv3x1_2", "conv1x3_2")]
        finally:
            ops.USE_TC = True
    xo = x0.clone().requires_grad_(True)
    sdo = {k: v.clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in sd.items()}
    yo = O.non_bottleneck_1d(xo, sdo, "blk", d, dropprob=p, training=True, drop_scale=None if drop is None else drop.view(n, c, 1, 1),
                             use_batchnorm=False)
    yo.backward(gy)
    ref = [yo.detach(), xo.grad] + [sdo[f"blk.{cn}.weight"].grad for cn in ("conv3x1_1", "conv1x3_2")] + \
          [sdo[f"blk.{cn}.bias"].grad for cn in ("conv3x1_1", "conv1x3_1", "conv3x1_2", "conv1x3_2")]
    labels = ["y", "dx", "dw1", "dw4", "db1", "db2", "db3", "db4"]
    for lab, a, b, r in zip(labels, res["tc"], res["generic"], ref):
        assert_rel_l2(a.numpy(), b.numpy(), 2e-2, f"tc vs generic {lab}")
        assert_rel_l2(a.numpy(), r.numpy(), 0.15, f"tc vs oracle {lab}")  # bf16 mask-flip noise, DESIGN.md section 2


def test_erfnet_nobn_net(golden_dir):
    """eval/erfnet_nobn.ERFNet(20) (the net eval_forwardTime.py times): fp32 check mode vs the reference's logits, bf16 within the gate."""
    from tests.nobn_cases import check_erfnet_nobn
    check_erfnet_nobn(golden_dir, "cuda", "fp32", tol=1e-4)
    check_erfnet_nobn(golden_dir, "cuda", "bf16", tol=2e-2)
