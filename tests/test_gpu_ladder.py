"""GPU: the fork's ladder regression net (train/net_regression_ladder.py, BASELINE config 4) through the C ABI: fp32 check mode vs the
reference's golden outputs / loss / gradients, bf16 (tcgen05 kernels at a tensor-core-eligible size) vs the fp32 oracle, and one
host-driven training step with torch.optim.Adam like train/main.py."""
import numpy as np
import pytest
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input
from tests.ladder_cases import check_ladder_net, check_ladder_net_classes, make_net
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu


def test_ladder_net_fp32_golden(golden_dir):
    check_ladder_net(golden_dir, "cuda", "fp32", out_tol=1e-4, grad_tol=2e-3)


def test_ladder_net_softmax_classes_fp32_golden(golden_dir):
    check_ladder_net_classes(golden_dir, "cuda", "fp32", out_tol=1e-4, grad_tol=2e-3)


def test_ladder_net_bf16_tensor_cores():
    """1 x 3 x 128 x 1024: the 64/128-channel blocks, the ladder convs and the heads run on tcgen05 (W % 128 == 0 at 1/4 and 1/8)."""
    m, sd = make_net("reg", "cuda", "bf16")
    x = golden_input((1, 3, 128, 1024), 81)
    with torch.no_grad():
        ref = O.regression_ladder_forward(x, sd)
    m.eval()
    with torch.no_grad():
        y = m(x.cuda())
    assert_rel(y.cpu().numpy(), ref.numpy(), 3e-2, "ladder net bf16 eval vs oracle")
    # training step: masked L1 loss, autograd through the kernels, torch Adam on the module parameters (train/main.py:213, 379-381)
    m.train()
    opt = torch.optim.Adam(m.parameters(), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
    tgt = golden_input((1, 128, 1024), 82) * 2.0
    tgt[golden_input((1, 128, 1024), 83) < 0.7] = -1.0
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = EV.L1LossMasked()(m(x.cuda()), tgt.cuda())
        loss.backward()
        opt.step()
        losses.append(loss.item())
    assert all(np.isfinite(losses)) and losses[-1] < losses[0], losses
    assert m.encoder.output_conv.weight.grad is None          # unused in this graph, as in the reference


def test_32_channel_block_folded_onto_the_64_channel_kernels():
    """LadderBlock(16): its non_bottleneck_1d(32) runs as a [W/2, 64]-folded view on conv1d_tc<64> (block-structured weights).  Against the
    same block on the generic kernels (ops.FOLD_32 = False, same bf16 storage): eval output, training output, input and parameter gradients."""
    from erfnet_pytorch_b200 import erfnet_blocks as EB
    from erfnet_pytorch_b200 import ops
    torch.manual_seed(0)
    blk = EB.LadderBlock(16).cuda()
    blk.precision = "bf16"
    sd = O.make_state_dict({k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=61)
    blk.load_state_dict(sd)
    x = (golden_input((2, 16, 8, 512), 121) - 0.3).cuda()
    lad = (golden_input((2, 16, 8, 512), 122) - 0.3).cuda()
    gout = (golden_input((2, 16, 8, 512), 123) - 0.5).cuda()
    res = {}
    for fold in (True, False):
        ops.FOLD_32 = fold
        try:
            l0 = ops.LAUNCHES
            blk.eval()
            with torch.no_grad():
                ye = blk(x, lad)
            blk.train()
            for p in blk.parameters():
                p.grad = None
            xi, li = x.clone().requires_grad_(True), lad.clone().requires_grad_(True)
            yt = blk(xi, li)
            yt.backward(gout)
            res[fold] = dict(ye=ye, yt=yt.detach(), dx=xi.grad, dl=li.grad, dw=blk.layer.conv1x3_1.weight.grad.clone(), dw3=blk.layer.conv3x1_2.weight.grad.clone(),
                             db=blk.layer.conv1x3_2.bias.grad.clone(), dwc=blk.conv.weight.grad.clone())
        finally:
            ops.FOLD_32 = True
    for k in res[True]:
        a, b = res[True][k].float(), res[False][k].float()
        assert a.shape == b.shape
        assert_rel(a.cpu().numpy(), b.cpu().numpy(), 3e-2, "folded vs generic: " + k)
    assert float((res[True]["ye"] - res[False]["ye"]).abs().max()) > 0      # really a different kernel path
