"""GPU: the generic tcgen05 conv (software im2col) vs fp32 torch convs of the same bf16 operands, for every shape class
it serves in ERFNet: stride-2 3x3 convs, stride-2 transposed convs (3x3 and 2x2), 1x1, and their data gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from erfnet_pytorch_b200 import ops
from erfnet_pytorch_b200.ops import ConvSpec
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu

CASES = [
    ("down64_128", ConvSpec(64, 64, 3, 3, 2, 1, 1), (2, 24, 40)),
    ("down16_64", ConvSpec(16, 48, 3, 3, 2, 1, 1), (2, 20, 36)),
    ("up128_64", ConvSpec(128, 64, 3, 3, 2, 1, 1, transposed=True, out_pad=1), (2, 9, 20)),
    ("up64_16", ConvSpec(64, 16, 3, 3, 2, 1, 1, transposed=True, out_pad=1), (3, 10, 24)),
    ("head16_20", ConvSpec(16, 20, 2, 2, 2, transposed=True), (2, 12, 40)),
    ("enc_head_1x1", ConvSpec(128, 20, 1, 1), (2, 8, 24)),
]


# row-aligned enumeration grids (width % 128 == 0): the TMA-loaded form of the kernel; the small sizes above keep the cp.async gather covered
CASES += [(c[0] + "_tma", c[1], (2, 4, 128 if c[1].transposed else 256)) for c in CASES[:5]] + [("enc_head_1x1_tma", CASES[5][1], (1, 3, 256))]


# ragged rows (round 2): any enumerated width >= 16 runs on the TMA loaders as ceil(EW / 128) row units, the last one zero-filled / clipped;
# narrower grids (EW < 16) keep the cp.async gather covered
CASES += [(c[0] + "_ragged", c[1], (2, 5, 100 if c[1].transposed else 200)) for c in CASES[:5]] + [("enc_head_1x1_ragged", CASES[5][1], (2, 3, 160))]
CASES += [(c[0] + "_gather", c[1], (2, 9, 12 if c[1].transposed else 24)) for c in CASES[:5]] + [("enc_head_1x1_gather", CASES[5][1], (3, 7, 12))]


def _torch_fwd(x_nchw, w, b, spec):
    if spec.transposed:
        return F.conv_transpose2d(x_nchw, w, b, stride=spec.stride, padding=(spec.pad_h, spec.pad_w), output_padding=spec.out_pad)
    return F.conv2d(x_nchw, w, b, stride=spec.stride, padding=(spec.pad_h, spec.pad_w))


@pytest.mark.parametrize("name,spec,nhw", CASES, ids=[c[0] for c in CASES])
def test_forward_and_data_gradient(name, spec, nhw):
    gen = torch.Generator().manual_seed(len(name))
    n, h, w = nhw
    wshape = (spec.cin, spec.cout, spec.kh, spec.kw) if spec.transposed else (spec.cout, spec.cin, spec.kh, spec.kw)
    wt = torch.randn(*wshape, generator=gen) / (spec.cin * spec.kh * spec.kw) ** 0.5
    bias = torch.randn(spec.cout, generator=gen) * 0.1
    x = torch.randn(n, h, w, spec.cin, generator=gen).bfloat16()
    wq = wt.bfloat16().float()
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = _torch_fwd(xr, wq, bias, spec)
    # forward (+ bias, relu) into a wider tensor at a channel offset, like the downsampler concat
    assert ops.im2col_supported(spec, torch.bfloat16)
    plan = ops.im2col_plan(spec, row_parity=False)   # one launch per output parity class: any output pitch / channel offset
    imgs = ops.im2col_pack(wt.cuda(), spec, False, plan)
    ho, wo = ref.shape[2:]
    c8 = ops.round_up(spec.cout, 8)
    out = torch.full((n, ho, wo, c8 + 8), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.conv_im2col_tc(x.cuda(), imgs, plan, spec, bias.cuda(), relu=True, out=out, out_coff=0)
    got = out[..., :spec.cout].float().cpu()
    assert_rel(got.numpy(), ref.relu().permute(0, 2, 3, 1).detach().numpy(), 1e-2, name + " fwd")
    assert torch.all(out[..., c8:] == 7.0), "channels outside the slice must be untouched"
    if spec.transposed and spec.stride == 2:
        # row-parity form (the default plan): column parities folded into N, whole output rows per launch, dense output
        plan2 = ops.im2col_plan(spec)
        assert plan2[0].get("merged") == c8 and len(plan2) < len(plan) + 1
        out2 = ops.conv_im2col_tc(x.cuda(), ops.im2col_pack(wt.cuda(), spec, False, plan2), plan2, spec, bias.cuda(), relu=True)
        assert_rel(out2[..., :spec.cout].float().cpu().numpy(), ref.relu().permute(0, 2, 3, 1).detach().numpy(), 1e-2, name + " fwd row-parity")
        assert_rel(out2[..., :spec.cout].float().cpu().numpy(), got.numpy(), 4e-3, name + " row-parity vs per-class")  # tap order differs
        with pytest.raises(RuntimeError):
            ops.conv_im2col_tc(x.cuda(), imgs, plan2, spec, bias.cuda(), out=out)   # needs a dense output
    # data gradient (only where the gather's input channels are a multiple of 16)
    dspec = ops.dgrad_spec(spec)
    if ops.im2col_supported(dspec, torch.bfloat16):
        dy = torch.randn(n, ho, wo, spec.cout, generator=gen).bfloat16()
        (gref,) = torch.autograd.grad(ref, xr, dy.float().permute(0, 3, 1, 2))
        dplan = ops.im2col_plan(dspec)
        dimgs = ops.im2col_pack(wt.cuda(), spec, True, dplan)
        dx = ops.conv_im2col_tc(dy.cuda(), dimgs, dplan, dspec, out_hw=(h, w))
        assert_rel(dx[..., :spec.cin].float().cpu().numpy(), gref.permute(0, 2, 3, 1).numpy(), 1e-2, name + " dgrad")


WG_CASES = CASES[:5] + [("stem3_13", ConvSpec(3, 13, 3, 3, 2, 1, 1), (2, 20, 44))]
# row-aligned dense grids (width % 64 == 0): the TMA-loaded form of the kernel (strided tensor maps); the sizes above keep the cp.async gather covered
WG_CASES += [(c[0] + "_tma", c[1], (2, 6, 64 if c[1].transposed else 128)) for c in CASES[:5]] + [("up128_64_tma_wide", CASES[2][1], (1, 3, 192))]


# ragged rows: ceil(EW / 64) units per row, the last one zero-filled; EW < 16 keeps the cp.async gather covered
WG_CASES += [(c[0] + "_ragged", c[1], (2, 5, 100 if c[1].transposed else 200)) for c in CASES[:5]]
WG_CASES += [(c[0] + "_gather", c[1], (2, 9, 12 if c[1].transposed else 24)) for c in CASES[:5]]


@pytest.mark.parametrize("name,spec,nhw", WG_CASES, ids=[c[0] for c in WG_CASES])
def test_weight_gradient(name, spec, nhw):
    gen = torch.Generator().manual_seed(3 + len(name))
    n, h, w = nhw
    xld = ops.round_up(spec.cin, 4)
    wshape = (spec.cin, spec.cout, spec.kh, spec.kw) if spec.transposed else (spec.cout, spec.cin, spec.kh, spec.kw)
    x = torch.zeros(n, h, w, xld).bfloat16()
    x[..., :spec.cin] = torch.randn(n, h, w, spec.cin, generator=gen).bfloat16()
    wt = torch.zeros(*wshape, requires_grad=True)
    y = _torch_fwd(x[..., :spec.cin].float().permute(0, 3, 1, 2), wt, None, spec)
    ho, wo = y.shape[2:]
    # dy lives in a wider tensor (the downsampler's concat gradient): extra channels hold unrelated data
    dyld = ops.round_up(spec.cout, 8) + 8
    dy = torch.randn(n, ho, wo, dyld, generator=gen).bfloat16()
    (ref,) = torch.autograd.grad(y, wt, dy[..., :spec.cout].float().permute(0, 3, 1, 2))
    assert ops.wgrad_im2col_supported(spec, torch.bfloat16)
    dw = torch.full(wshape, 0.5, device="cuda")   # accumulates into existing contents
    ops.conv_wgrad_im2col_tc(x.cuda(), dy.cuda(), spec, dw_out=dw)
    assert_rel((dw.cpu() - 0.5).numpy(), ref.numpy(), 3e-3, name + " dW")


def test_head_data_gradient_row_folded():
    """ConvTranspose2d(16, 20, 2, stride 2) data gradient through the row-folded view (K = 2 x 48) vs torch autograd."""
    gen = torch.Generator().manual_seed(11)
    spec = ConvSpec(16, 20, 2, 2, 2, 0, 0, transposed=True)
    n, h, w = 2, 12, 72
    wt = torch.randn(16, 20, 2, 2, generator=gen) * 0.2
    x = torch.randn(n, 16, h, w, generator=gen, requires_grad=True)
    y = torch.nn.functional.conv_transpose2d(x, wt.bfloat16().float(), stride=2)
    dy = torch.zeros(n, 2 * h, 2 * w, 24).bfloat16()
    dy[..., :20] = torch.randn(n, 2 * h, 2 * w, 20, generator=gen).bfloat16()
    dy[..., 20:] = 3.0          # pad channels hold junk: the zero weight rows must cancel it
    (gref,) = torch.autograd.grad(y, x, dy[..., :20].float().permute(0, 3, 1, 2))
    assert ops.convT2x2_dgrad_tc_supported(spec, dy.cuda())
    dx = ops.convT2x2_dgrad_tc(dy.cuda(), wt.cuda(), spec)
    assert_rel(dx.float().cpu().numpy(), gref.permute(0, 2, 3, 1).numpy(), 1e-2, "head dgrad")


def test_channel_sum_padded_24():
    x = torch.randn(3, 16, 40, 24).bfloat16().cuda()
    out = torch.zeros(24, device="cuda")
    ops.channel_sum_slice(x, 24, out)
    np.testing.assert_allclose(out.cpu().numpy(), x.float().sum((0, 1, 2)).cpu().numpy(), rtol=2e-3, atol=2e-2)


@pytest.mark.parametrize("shape", [(2, 3, 64, 256), (1, 3, 40, 72), (3, 3, 8, 8)])
def test_fused_nchw_head_matches_the_two_pass_form(shape):
    """Inference: the decoder head writes the NCHW fp32 logits from its epilogue.  Against the NHWC bf16 head + layout kernel it differs only by
    the bf16 rounding of the logits it no longer performs; against the fp32 check mode it is within the bf16 gate."""
    from erfnet_pytorch_b200 import erfnet as E
    from oracle import erfnet_oracle as O
    m = E.ERFNet(20)
    m.load_state_dict(O.make_state_dict(O.erfnet_param_shapes(20), 3))
    m = m.cuda().eval()
    x = torch.rand(shape, generator=torch.Generator().manual_seed(12)).cuda()
    with torch.no_grad():
        m(x)                                                               # first call packs the eval-mode weights
        l0 = ops.LAUNCHES
        fused = m(x)
    n_fused = ops.LAUNCHES - l0
    ops.FUSE_NCHW_HEAD = False
    try:
        with torch.no_grad():
            m(x)
            l0 = ops.LAUNCHES
            two_pass = m(x)
        n_two = ops.LAUNCHES - l0
    finally:
        ops.FUSE_NCHW_HEAD = True
    assert fused.shape == two_pass.shape == (shape[0], 20, shape[2], shape[3]) and fused.dtype == torch.float32
    assert n_fused == n_two - 1                                            # the layout pass is gone
    scale = two_pass.abs().max().item()
    assert (fused - two_pass).abs().max().item() <= 2.0 ** -8 * scale      # one bf16 rounding of the logits
    assert (fused - two_pass).abs().max().item() > 0                       # ... and it really is a different (more exact) path
    m.precision = "fp32"
    with torch.no_grad():
        ref = m(x)
    assert (fused - ref).abs().max().item() <= 2e-2 * ref.abs().max().item()


@pytest.mark.parametrize("dyld,w", [(24, 72), (32, 72), (24, 128)])
def test_head_weight_gradient_column_folded(dyld, w):
    """ConvTranspose2d(16, 20, 2, stride 2) weight gradient through the column-folded view of dy (two taps of 2 * ld channels) vs torch
    autograd and vs the four-tap form it replaces."""
    gen = torch.Generator().manual_seed(77)
    spec = ConvSpec(16, 20, 2, 2, 2, 0, 0, transposed=True)
    n, h = 3, 24
    x = torch.randn(n, h, w, 16, generator=gen).bfloat16()
    wt = torch.zeros(16, 20, 2, 2, requires_grad=True)
    y = _torch_fwd(x.float().permute(0, 3, 1, 2), wt, None, spec)
    dy = torch.zeros(n, 2 * h, 2 * w, dyld).bfloat16()
    dy[..., :20] = torch.randn(n, 2 * h, 2 * w, 20, generator=gen).bfloat16()
    if dyld > 24:
        dy[..., 24:] = 3.0    # channels beyond the padded logits belong to somebody else: must not leak into the gradient
    (ref,) = torch.autograd.grad(y, wt, dy[..., :20].float().permute(0, 3, 1, 2))
    got = []
    for fold in (True, False):
        ops.FOLD_HEAD_WGRAD = fold
        try:
            dw = torch.full((16, 20, 2, 2), -0.25, device="cuda")
            ops.conv_wgrad_im2col_tc(x.cuda(), dy.cuda(), spec, dw_out=dw)
            got.append((dw.cpu() + 0.25).numpy())
        finally:
            ops.FOLD_HEAD_WGRAD = True
    assert_rel(got[0], ref.numpy(), 3e-3, "folded head dW")
    assert_rel(got[0], got[1], 1e-3, "folded vs four-tap head dW")
