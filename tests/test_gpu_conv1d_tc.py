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


# widths: multiples of 128 (whole units) and ragged ones -- the reference fork's own 640 x 480 frames give 320 / 160 / 80, the 224 x 224 ImageNet net 56 / 28
@pytest.mark.parametrize("c,w", [(64, 256), (128, 128), (64, 128), (128, 384), (16, 512), (16, 128), (64, 160), (128, 80), (16, 320), (128, 28), (64, 56), (16, 200)])
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
    # fused epilogue reductions: BN statistics and column sums of the tensor as stored
    stats = torch.zeros(2 * c, device="cuda", dtype=torch.float64)
    colsum = torch.zeros(c, device="cuda", dtype=torch.float32)
    y = ops.conv1d_tc(xc, img, bc, horizontal, dil, stats=stats, colsum=colsum)
    yf = y.double().cpu()
    np.testing.assert_allclose(stats[:c].cpu().numpy(), yf.sum((0, 1, 2)).numpy(), rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(stats[c:].cpu().numpy(), (yf ** 2).sum((0, 1, 2)).numpy(), rtol=1e-4)
    np.testing.assert_allclose(colsum.cpu().numpy(), yf.sum((0, 1, 2)).numpy(), rtol=1e-4, atol=1e-2)
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
        # all 17 blocks x 4 convs on the tensor cores: the 64- and 16-channel blocks as fused pairs, the 128-channel ones as single convs
        n_single, n_pair = kt.get("conv1d_tc_kernel", {}).get("n", 0), kt.get("conv1d_pair_tc_kernel", {}).get("n", 0)
        assert n_single + 2 * n_pair == 68 and n_single >= 32, kt
        ops.USE_TC = False
        try:
            y_gen = m(x.cuda()).cpu()
        finally:
            ops.USE_TC = True
    assert_rel(y_tc.numpy(), ref.numpy(), 2e-2, "tcgen05 path vs oracle")   # north_star bf16 gate
    assert_rel(y_gen.numpy(), ref.numpy(), 2e-2, "generic path vs oracle")
    agree = (y_tc.argmax(1) == ref.argmax(1)).float().mean().item()
    print(f"argmax agreement tc-vs-oracle: {agree:.5f}; max rel err {np.abs(y_tc.numpy() - ref.numpy()).max() / np.abs(ref.numpy()).max():.3e}")


@pytest.mark.parametrize("c,d,hw", [(64, 1, (8, 256)), (128, 4, (12, 128)), (128, 16, (40, 128)), (16, 1, (10, 512))])
def test_block_train_fwd_bwd_on_tensor_cores(c, d, hw):
    """Train-mode non_bottleneck_1d (batch-stat BN, fixed dropout mask): forward, dx, dW, dgamma with the tcgen05 convs."""
    blk = E.non_bottleneck_1d(c, 0.1, d)
    bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=3)
    blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
    blk.cuda().train()
    x = golden_input((2, c) + hw, 40 + d) - 0.3
    gy = golden_input((2, c) + hw, 41) - 0.5
    keep = (torch.rand(2, c, generator=torch.Generator().manual_seed(1)) > 0.1).float() / 0.9
    # oracle (fp32 CPU)
    sdo = {k: v.clone() for k, v in bsd.items()}
    for k, v in sdo.items():
        if v.is_floating_point() and "running" not in k:
            v.requires_grad_(True)
    xo = x.clone().requires_grad_(True)
    yo = O.non_bottleneck_1d(xo, sdo, "blk", d, 0.1, training=True, drop_scale=keep.view(2, c, 1, 1))
    yo.backward(gy)
    # ours, bf16, through the NHWC internals so that the dropout mask can be injected: generic CUDA-core convs, then tcgen05
    def l2(a, b):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        return float(np.sqrt(((a - b) ** 2).sum() / (b ** 2).sum()))

    runs = {}
    for use_tc in (False, True):
        blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
        blk.zero_grad()
        ops.USE_TC = use_tc
        try:
            ctx = E._Ctx("bf16", True, {blk: keep})
            xn = ops.nchw_to_nhwc(x.cuda(), c, torch.bfloat16).requires_grad_(True)
            ops.start_kernel_timing()
            yn = E._run(blk, xn, ctx)
            yn.backward(ops.nchw_to_nhwc(gy.cuda(), c, torch.bfloat16))
            kt = ops.stop_kernel_timing()
        finally:
            ops.USE_TC = True
        assert kt.get("conv1d_tc_kernel", {}).get("n", 0) == (8 if use_tc else 0), kt  # 4 forward convs + 4 data gradients on tcgen05
        runs[use_tc] = dict(y=ops.nhwc_to_nchw(yn.detach(), c).cpu().numpy(), dx=ops.nhwc_to_nchw(xn.grad, c).cpu().numpy(),
                            dw4=blk.conv1x3_2.weight.grad.cpu().numpy(), dw1=blk.conv3x1_1.weight.grad.cpu().numpy(),
                            dg1=blk.bn1.weight.grad.cpu().numpy(), rv2=blk.bn2.running_var.cpu().numpy())
    ref = dict(y=yo.detach().numpy(), dx=xo.grad.numpy(), dw4=sdo["blk.conv1x3_2.weight"].grad.numpy(), dw1=sdo["blk.conv3x1_1.weight"].grad.numpy(),
               dg1=sdo["blk.bn1.weight"].grad.numpy(), rv2=sdo["blk.bn2.running_var"].numpy())
    # (a) tcgen05 vs generic kernels on identical bf16 data: the kernels compute the same thing
    for k in ref:
        assert l2(runs[True][k], runs[False][k]) < 6e-2, (k, l2(runs[True][k], runs[False][k]))
    # (b) vs the fp32 oracle: forward within the bf16 gate; gradients loosely -- bf16 activations flip ~0.5 % of the ReLU
    # masks near zero, which alone is ~sqrt(0.005) = 7 % relative L2 on a gradient (the fp32 mode holds 1e-6 on the same test)
    assert_rel(runs[True]["y"], ref["y"], 3e-2, "train fwd")
    assert_rel(runs[True]["rv2"], ref["rv2"], 2e-2, "running_var2")
    for k in ("dx", "dw4", "dw1", "dg1"):
        assert l2(runs[True][k], ref[k]) < 0.2, (k, l2(runs[True][k], ref[k]))


@pytest.mark.parametrize("c,w", [(64, 256), (128, 128), (64, 64), (128, 192), (16, 512), (16, 64), (64, 160), (128, 80), (16, 320), (128, 28), (64, 56), (16, 200)])
@pytest.mark.parametrize("horizontal", [False, True])
@pytest.mark.parametrize("dil", [1, 2, 16])
def test_conv1d_wgrad_tc_matches_autograd(c, w, horizontal, dil):
    gen = torch.Generator().manual_seed(c * 3 + w + dil)
    n, h = 3, 37
    x = torch.randn(n, h, w, c, generator=gen).bfloat16()
    dy = torch.randn(n, h, w, c, generator=gen).bfloat16()
    wt = torch.zeros(c, c, *((1, 3) if horizontal else (3, 1)), requires_grad=True)
    pad, dl = ((0, dil), (1, dil)) if horizontal else ((dil, 0), (dil, 1))
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, padding=pad, dilation=dl)
    (ref,) = torch.autograd.grad(y, wt, dy.float().permute(0, 3, 1, 2))
    dw, db = ops.conv1d_wgrad_tc(x.cuda(), dy.cuda(), horizontal, dil)
    assert dw.shape == wt.shape
    assert_rel(dw.cpu().numpy(), ref.numpy(), 2e-3, "dW")
    assert_rel(db.cpu().numpy(), dy.float().sum((0, 1, 2)).numpy(), 2e-3, "db")


@pytest.mark.parametrize("c,n,h,w", [(128, 6, 64, 128), (64, 6, 128, 256), (16, 6, 256, 512), (64, 1, 3, 64), (128, 2, 5, 192)])
def test_conv1d_wgrad_reduction_forms_agree(c, n, h, w):
    """The three reduction forms of the weight gradient -- bulk (default: every CTA adds its accumulators to the gradient with cp.reduce.async.bulk,
    one launch), in-kernel (grid barrier, every CTA sums its slice of the partial-sum slabs) and the separate reduce kernel -- add the same partial
    sums in different orders, ACCUMULATED into existing gradient contents; bench shapes and grids smaller than the SM count."""
    gen = torch.Generator().manual_seed(c + n + h)
    x = torch.randn(n, h, w, c, generator=gen).bfloat16().cuda()
    dy = torch.randn(n, h, w, c, generator=gen).bfloat16().cuda()
    got = []
    try:
        for bulk, fused, launches in ((True, False, 2), (False, True, 2), (False, False, 4)):
            ops.set_wgrad_bulk(bulk)
            ops.WGRAD_FUSED_REDUCE = fused
            dw = torch.full((c, c, 3, 1), 0.125, device="cuda")
            before = ops.LAUNCHES
            for _ in range(2):                                   # two accumulating calls
                ops.conv1d_wgrad_tc(x, dy, False, 2, dw_out=dw, want_bias=False)
            assert ops.LAUNCHES - before == launches
            got.append(dw.cpu().numpy() - 0.125)
    finally:
        ops.WGRAD_FUSED_REDUCE = False
        ops.set_wgrad_bulk(True)
    assert_rel(got[0], got[2], 2e-5, "bulk vs two-launch dW")
    assert_rel(got[1], got[2], 2e-5, "fused vs two-launch dW")
    assert np.abs(got[0]).max() > 1.0


def test_conv1d_wgrad_bulk_needs_aligned_gradient():
    """A gradient view that is not 16-byte aligned cannot be the target of a bulk reduction: the library falls back to the reduce kernel."""
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 8, 128, 64, generator=gen).bfloat16().cuda()
    dy = torch.randn(2, 8, 128, 64, generator=gen).bfloat16().cuda()
    ref, _ = ops.conv1d_wgrad_tc(x, dy, True, 1, want_bias=False)
    buf = torch.zeros(64 * 64 * 3 + 4, device="cuda")
    dw = buf[1:1 + 64 * 64 * 3].view(64, 64, 1, 3)
    assert dw.data_ptr() % 16 != 0
    before = ops.LAUNCHES
    ops.conv1d_wgrad_tc(x, dy, True, 1, dw_out=dw, want_bias=False)
    assert ops.LAUNCHES - before == 2
    assert_rel(dw.cpu().numpy(), ref.cpu().numpy(), 2e-5, "unaligned dW")
    assert float(buf[0]) == 0.0 and float(buf[-1]) == 0.0


@pytest.mark.parametrize("c,n,h,w,dil", [(128, 6, 64, 128, 8), (64, 6, 128, 256, 1), (16, 6, 256, 512, 1), (64, 1, 3, 64, 2), (128, 2, 5, 192, 16)])
def test_conv1d_wgrad_group_matches_single_launches(c, n, h, w, dil):
    """The grouped weight-gradient kernel (four jobs of one block in one launch: 3x1, 1x3, dilated 3x1, dilated 1x3, each with its own x / dy)
    against four single launches: identical partial sums, ACCUMULATED into existing gradient contents."""
    gen = torch.Generator().manual_seed(c + n + h + dil)
    xs = [torch.randn(n, h, w, c, generator=gen).bfloat16().cuda() for _ in range(4)]
    dys = [torch.randn(n, h, w, c, generator=gen).bfloat16().cuda() for _ in range(4)]
    geo = [(False, 1), (True, 1), (False, dil), (True, dil)]
    ref = []
    for (hz, d), x, dy in zip(geo, xs, dys):
        dw = torch.full((c, c, 1, 3) if hz else (c, c, 3, 1), 0.25, device="cuda")
        ops.conv1d_wgrad_tc(x, dy, hz, d, dw_out=dw, want_bias=False)
        ref.append(dw)
    got = [torch.full_like(r, 0.25) for r in ref]
    before = ops.LAUNCHES
    ops.conv1d_wgrad_tc_group([(x, dy, hz, d, dw) for (hz, d), x, dy, dw in zip(geo, xs, dys, got)])
    assert ops.LAUNCHES - before == 2
    for a, b, (hz, d) in zip(got, ref, geo):
        assert_rel((a - 0.25).cpu().numpy(), (b - 0.25).cpu().numpy(), 2e-5, f"group vs single, horizontal={hz} dil={d}")
    # a shorter job list (the kernel takes 1..4)
    one = torch.zeros_like(ref[2])
    ops.conv1d_wgrad_tc_group([(xs[2], dys[2], False, dil, one)])
    assert_rel(one.cpu().numpy(), (ref[2] - 0.25).cpu().numpy(), 2e-5, "single-job group")


def test_pack_kernel_matches_host_packing():
    for c, shape in ((64, (64, 64, 3, 1)), (128, (128, 128, 1, 3)), (16, (16, 16, 3, 1))):
        w = torch.randn(*shape, generator=torch.Generator().manual_seed(c)).cuda()
        f, d = ops.pack_conv1d_tc_pair(w)
        assert torch.equal(f.view(-1), ops.pack_conv1d_tc(w).view(-1))
        assert torch.equal(d.view(-1), ops.pack_conv1d_tc(w, dgrad=True).view(-1))
