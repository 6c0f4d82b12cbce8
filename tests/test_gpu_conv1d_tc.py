"""GPU: the tcgen05 conv1d kernel (erf_conv1d_tc) vs an fp32 torch conv of the same bf16-rounded operands, and the
whole net with the tensor-core path on vs the CPU oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from erfnet_pytorch_b200 import erfnet as E
from erfnet_pytorch_b200 import ops
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("c,w", [(64, 256), (128, 128), (64, 128), (128, 384)])
@pytest.mark.parametrize("horizontal", [False, True])
@pytest.mark.parametrize("dil", [1, 2, 16])
def test_conv1d_tc_matches_fp32_conv(c, w, horizontal, dil):
    gen = torch.Generator().manual_seed(c + w + dil)
    n, h = 2, 20
    x = (torch.randn(n, h, w, c, generator=gen)).bfloat16()
    wt = (torch.randn(c, c, *((1, 3) if horizontal else (3, 1)), generator=gen) / (3 * c) ** 0.5)
    bias = torch.randn(c, generator=gen) * 0.1
    res = torch.randn(n, h, w, c, generator=gen).bfloat16()
    mask = torch.randn(n, h, w, c, generator=gen).bfloat16()
    wq = wt.bfloat16().float()
    pad, dl = ((0, dil), (1, dil)) if horizontal else ((dil, 0), (dil, 1))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wq, bias, padding=pad, dilation=dl).permute(0, 2, 3, 1)
    img = ops.pack_conv1d_tc(wt.cuda())
    xc, bc = x.cuda(), bias.cuda()
    # plain conv + bias
    y = ops.conv1d_tc(xc, img, bc, horizontal, dil)
    assert_rel(y.float().cpu().numpy(), ref.numpy(), 1e-2, "conv+bias")
    # residual + relu
    y = ops.conv1d_tc(xc, img, bc, horizontal, dil, relu=True, res=res.cuda())
    assert_rel(y.float().cpu().numpy(), (ref + res.float()).relu().numpy(), 1e-2, "res+relu")
    # relu-backward mask, no bias
    y = ops.conv1d_tc(xc, img, None, horizontal, dil, mask=mask.cuda())
    ref_m = F.conv2d(x.float().permute(0, 3, 1, 2), wq, None, padding=pad, dilation=dl).permute(0, 2, 3, 1) * (mask.float() > 0)
    assert_rel(y.float().cpu().numpy(), ref_m.numpy(), 1e-2, "mask")
    # data-gradient packing: conv_transpose == conv with flipped taps / transposed channels
    dimg = ops.pack_conv1d_tc(wt.cuda(), dgrad=True)
    g = ops.conv1d_tc(xc, dimg, None, horizontal, dil)
    ref_g = F.conv_transpose2d(x.float().permute(0, 3, 1, 2), wq, None, padding=pad, dilation=dl).permute(0, 2, 3, 1)
    assert_rel(g.float().cpu().numpy(), ref_g.numpy(), 1e-2, "dgrad")


def test_whole_net_with_tensor_core_blocks():
    """64 x 1024 input: every non_bottleneck_1d of the encoder/decoder at 64 and 128 channels runs on tcgen05."""
    m = E.ERFNet(20)
    sd = O.make_state_dict(O.erfnet_param_shapes(20), 4)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = golden_input((2, 3, 64, 1024), 11)
    with torch.no_grad():
        ref = O.erfnet_forward(x, sd)
        ops.start_kernel_timing()
        y_tc = m(x.cuda()).cpu()
        kt = ops.stop_kernel_timing()
        assert kt.get("conv1d_tc_kernel", {}).get("n", 0) == 60, kt  # 15 blocks x 4 convs on the tensor cores
        ops.USE_TC = False
        try:
            y_gen = m(x.cuda()).cpu()
        finally:
            ops.USE_TC = True
    assert_rel(y_tc.numpy(), ref.numpy(), 2e-2, "tcgen05 path vs oracle")   # north_star bf16 gate
    assert_rel(y_gen.numpy(), ref.numpy(), 2e-2, "generic path vs oracle")
    agree = (y_tc.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"argmax agreement tc-vs-oracle: {agree:.5f}; max rel err {np.abs(y_tc.numpy() - ref.numpy()).max() / np.abs(ref.numpy()).max():.3e}")
