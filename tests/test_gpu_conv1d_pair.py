"""GPU: fused 3x1 -> ReLU -> 1x3 pair kernel (tcgen05, intermediate row in shared memory) vs two conv1d_tc launches (bit-identical)
and vs an fp32 torch reference of the same bf16 operands."""
import pytest
import torch
import torch.nn.functional as F

from erfnet_pytorch_b200 import ops
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu

CASES = [(64, 256, 1, 2, 9), (64, 128, 1, 1, 5), (64, 256, 2, 3, 7), (16, 512, 1, 2, 6), (16, 128, 4, 1, 11), (64, 256, 1, 6, 128)]


@pytest.mark.parametrize("c,w,dil,n,h", CASES)
@pytest.mark.parametrize("with_res", [False, True])
def test_pair_matches_two_launches(c, w, dil, n, h, with_res):
    gen = torch.Generator().manual_seed(c + w + dil)
    x = torch.randn(n, h, w, c, generator=gen).bfloat16().cuda()
    w1 = torch.randn(c, c, 3, 1, generator=gen) / (3 * c) ** 0.5
    w2 = torch.randn(c, c, 1, 3, generator=gen) / (3 * c) ** 0.5
    b1, b2 = (torch.randn(c, generator=gen) * 0.1).cuda(), (torch.randn(c, generator=gen) * 0.1).cuda()
    res = torch.randn(n, h, w, c, generator=gen).bfloat16().cuda() if with_res else None
    i1, i2 = ops.pack_conv1d_tc(w1.cuda()), ops.pack_conv1d_tc(w2.cuda())
    assert ops.conv1d_pair_tc_supported(c, w, dil, torch.bfloat16)
    t = ops.conv1d_tc(x, i1, b1, False, dil, relu=True)
    ref2 = ops.conv1d_tc(t, i2, b2, True, dil, relu=True, res=res)
    got = ops.conv1d_pair_tc(x, i1, b1, i2, b2, dil, relu=True, res=res)
    torch.cuda.synchronize()
    assert torch.equal(got, ref2), f"max diff {(got.float() - ref2.float()).abs().max().item()}"
    # and against torch on the same bf16 operands (intermediate rounded to bf16 like the stored tensor)
    xf = x.float().permute(0, 3, 1, 2).cpu()
    t_ref = F.relu(F.conv2d(xf, w1.bfloat16().float(), b1.cpu(), padding=(dil, 0), dilation=(dil, 1))).bfloat16().float()
    y_ref = F.conv2d(t_ref, w2.bfloat16().float(), b2.cpu(), padding=(0, dil), dilation=(1, dil))
    if with_res:
        y_ref = y_ref + res.float().permute(0, 3, 1, 2).cpu()
    assert_rel(got.float().permute(0, 3, 1, 2).cpu().numpy(), F.relu(y_ref).numpy(), 1e-2, "pair vs torch")


def test_pair_unsupported_shapes():
    assert not ops.conv1d_pair_tc_supported(128, 128, 1, torch.bfloat16)   # two 96 KB weight images do not fit
    assert not ops.conv1d_pair_tc_supported(64, 512, 1, torch.bfloat16)    # 4 tiles x 64 columns x 4 accumulator sets > 512 TMEM columns
    assert not ops.conv1d_pair_tc_supported(64, 200, 1, torch.bfloat16)
