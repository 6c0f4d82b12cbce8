"""GPU parity tests proper: the CUDA path through the C ABI vs the CPU oracle on the same seeded inputs.
fp32 "check mode" must hold 1e-4 max relative error (north_star); bf16 2e-2."""
import os

import numpy as np
import pytest
import torch

from erfnet_pytorch_b200 import erfnet as E
from erfnet_pytorch_b200 import ops
from erfnet_pytorch_b200.evaluation_functions import CrossEntropyLoss2d
from erfnet_pytorch_b200.iouEval import iouEval
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input, golden_labels
from tests.test_host_logic import WHOLE_NET_GRAD_TOL, assert_rel

pytestmark = pytest.mark.gpu
FP32_TOL = 1e-4   # north_star: logits within 1e-4 in the fp32 check mode
BF16_TOL = 2e-2   # north_star: logits within 2e-2 max relative error in bf16


def _model(precision, seed=0, nc=20):
    m = E.ERFNet(nc)
    m.precision = precision
    sd = O.make_state_dict(O.erfnet_param_shapes(nc), seed)
    m.load_state_dict(sd, strict=True)
    return m.cuda(), sd


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_whole_net_eval_vs_oracle_and_golden(precision, tol, golden_dir):
    g = np.load(os.path.join(golden_dir, "erfnet_eval_64x128.npz"))
    m, sd = _model(precision)
    m.eval()
    x = golden_input((1, 3, 64, 128), 0)
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        ye = m(x.cuda(), only_encode=True).cpu()
        ref = O.erfnet_forward(x, sd)
    assert y.shape == (1, 20, 64, 128) and y.dtype == torch.float32
    assert_rel(y.numpy(), g["logits"], tol, "logits vs golden")
    assert_rel(y.numpy(), ref.numpy(), tol, "logits vs oracle")
    assert_rel(ye.numpy(), g["enc_logits"], tol, "encoder logits vs golden")


@pytest.mark.parametrize("precision,tol", [("fp32", FP32_TOL), ("bf16", BF16_TOL)])
def test_larger_image_and_batch(precision, tol):
    """Non-square, batch 3, all dilations active in the interior (H/8 = 24 > 16)."""
    m, sd = _model(precision, seed=2)
    m.eval()
    x = golden_input((3, 3, 192, 320), 3)
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        ref = O.erfnet_forward(x, sd)
    assert_rel(y.numpy(), ref.numpy(), tol)
    if precision == "fp32":
        assert (y.argmax(1) == ref.argmax(1)).float().mean() > 0.999


def test_blocks_train_fwd_bwd_fp32(golden_dir):
    g = np.load(os.path.join(golden_dir, "blocks.npz"))
    cases = {"nb64_d1": (64, 1, (8, 24)), "nb128_d2": (128, 2, (8, 12)), "nb128_d16": (128, 16, (18, 20)),
             "nb16_d1": (16, 1, (16, 24)), "nb128_d4": (128, 4, (8, 16)), "nb128_d8": (128, 8, (10, 18))}
    for name, (c, d, hw) in cases.items():
        blk = E.non_bottleneck_1d(c, 0.0, d)
        blk.precision = "fp32"
        bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=3)
        blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
        blk.cuda()
        x = (golden_input((2, c) + hw, 10 + d) - 0.3).cuda()
        blk.eval()
        with torch.no_grad():
            assert_rel(blk(x).cpu().numpy(), g[name + "_eval"], FP32_TOL, name + " eval")
        blk.train()
        x.requires_grad_(True)
        y = blk(x)
        y.backward((golden_input(tuple(y.shape), 77) - 0.5).cuda())
        assert_rel(y.detach().cpu().numpy(), g[name + "_train"], 2e-4, name + " train")
        assert_rel(x.grad.cpu().numpy(), g[name + "_train_dx"], 1e-3, name + " dx")
        assert_rel(blk.conv1x3_2.weight.grad.cpu().numpy(), g[name + "_train_dw_conv1x3_2"], 1e-3, name + " dw4")
        assert_rel(blk.conv3x1_1.weight.grad.cpu().numpy(), g[name + "_train_dw_conv3x1_1"], 1e-3, name + " dw1")
        assert_rel(blk.bn1.weight.grad.cpu().numpy(), g[name + "_train_dgamma1"], 1e-3, name + " dgamma1")
        assert_rel(blk.bn2.running_mean.cpu().numpy(), g[name + "_train_rm2"], 1e-4, name + " rm2")
        assert_rel(blk.bn2.running_var.cpu().numpy(), g[name + "_train_rv2"], 1e-4, name + " rv2")
    for name, (ci, co, hw) in {"down3_16": (3, 16, (16, 32)), "down16_64": (16, 64, (8, 16)), "down64_128": (64, 128, (8, 8))}.items():
        blk = E.DownsamplerBlock(ci, co)
        blk.precision = "fp32"
        bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=4)
        blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
        blk.cuda()
        x = (golden_input((2, ci) + hw, 20 + ci) - 0.3).cuda()
        blk.eval()
        with torch.no_grad():
            assert_rel(blk(x).cpu().numpy(), g[name + "_eval"], FP32_TOL, name + " eval")
        blk.train()
        x.requires_grad_(True)
        y = blk(x)
        y.backward((golden_input(tuple(y.shape), 78) - 0.5).cuda())
        assert_rel(y.detach().cpu().numpy(), g[name + "_train"], 2e-4, name + " train")
        assert_rel(x.grad.cpu().numpy(), g[name + "_train_dx"], 1e-3, name + " dx")
        assert_rel(blk.conv.weight.grad.cpu().numpy(), g[name + "_train_dw"], 1e-3, name + " dw")
    for name, (ci, co, hw) in {"up128_64": (128, 64, (4, 8)), "up64_16": (64, 16, (8, 8))}.items():
        blk = E.UpsamplerBlock(ci, co)
        blk.precision = "fp32"
        bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=5)
        blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
        blk.cuda()
        x = (golden_input((2, ci) + hw, 30 + ci) - 0.3).cuda()
        blk.eval()
        with torch.no_grad():
            assert_rel(blk(x).cpu().numpy(), g[name + "_eval"], FP32_TOL, name + " eval")
        blk.train()
        x.requires_grad_(True)
        y = blk(x)
        y.backward((golden_input(tuple(y.shape), 79) - 0.5).cuda())
        assert_rel(y.detach().cpu().numpy(), g[name + "_train"], 2e-4, name + " train")
        assert_rel(x.grad.cpu().numpy(), g[name + "_train_dx"], 1e-3, name + " dx")
        assert_rel(blk.conv.weight.grad.cpu().numpy(), g[name + "_train_dw"], 1e-3, name + " dw")


def test_train_step_vs_golden_fp32(golden_dir):
    g = np.load(os.path.join(golden_dir, "train_step_32x64.npz"))
    m, sd = _model("fp32")
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    m.train()
    x = golden_input((2, 3, 32, 64), 5).cuda()
    y = golden_labels((2, 32, 64), 5).cuda()
    out = m(x)
    loss = CrossEntropyLoss2d()(out, y)
    loss.backward()
    assert_rel(out.detach().cpu().numpy(), g["logits_step0"], 3e-3, "train logits")
    assert abs(float(loss) - g["losses"][0]) <= 1e-4 * g["losses"][0]
    params = dict(m.named_parameters())
    for k in g.files:
        if k.startswith("grad0/") and np.abs(g[k]).max() > 1e-5:
            assert_rel(params[k[6:]].grad.cpu().numpy(), g[k], WHOLE_NET_GRAD_TOL, k)
    assert sorted(k for k, p in params.items() if p.grad is None) == sorted(g["no_grad_keys"].tolist())


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-3), ("bf16", 6e-2)])
def test_train_step_vs_oracle_128x256(precision, tol):
    """Train-mode forward (batch-statistics BN with fixed Dropout2d masks), weighted CE and backward at a size where
    every BN sees >= 1024 samples (the 32x64 golden is too ill-conditioned for bf16: its BNs see 64 samples)."""
    m, sd = _model(precision, seed=1)
    m.train()
    x = golden_input((2, 3, 128, 256), 6)
    y = golden_labels((2, 128, 256), 6)
    gen = torch.Generator().manual_seed(3)
    masks, masks_o = {}, {}
    for name, mod in m.named_modules():
        if isinstance(mod, E.non_bottleneck_1d) and mod.dropout.p != 0:
            keep = (torch.rand(2, mod.chann, generator=gen) > 0.1).float() / 0.9
            masks[mod] = keep
            masks_o[name] = keep.view(2, mod.chann, 1, 1)
    wv = torch.linspace(0.5, 1.5, 20)
    logits = m.forward_nhwc(x.cuda(), drop_scales=masks)
    loss = CrossEntropyLoss2d(weight=wv)(logits, y.cuda(), nhwc_classes=20)
    loss.backward()
    sd_o = {k: v.clone() for k, v in sd.items()}
    po = {k: v.requires_grad_(True) for k, v in sd_o.items() if v.is_floating_point() and "running" not in k}
    out_o = O.erfnet_forward(x, sd_o, training=True, drop_scales=masks_o)
    lo = O.cross_entropy_2d(out_o, y, weight=wv)
    lo.backward()
    got = logits.detach()[..., :20].float().cpu().permute(0, 3, 1, 2).numpy()
    if precision == "fp32":
        assert_rel(got, out_o.detach().numpy(), tol, "train logits")
        assert abs(float(loss) - float(lo)) <= 1e-4 * float(lo)
    else:
        # Train-mode BN of this random-weight net amplifies rounding noise ~1.3x per block (tests/emu.py reproduces the same
        # growth on the CPU with bf16 storage: 0.6 % after the stem -> 60 % relative L2 at the logits), so in bf16 only the
        # loss is comparable here; per-kernel accuracy is pinned in fp32 mode and by the tcgen05-vs-generic block tests,
        # and the north_star's bf16 training gate is the 200-step loss curve (tests/test_gpu_training_curve.py).
        assert np.isfinite(got).all()
        assert abs(float(loss) - float(lo)) <= 5e-2 * float(lo)
    msd = m.state_dict()
    for k in ("encoder.layers.9.bn2.running_var", "decoder.layers.3.bn.running_mean", "encoder.initial_block.bn.running_var"):
        assert_rel(msd[k].cpu().numpy(), sd_o[k].numpy(), 1e-3 if precision == "fp32" else 0.2, k)
    if precision == "fp32":
        for k, p in m.named_parameters():
            if p.grad is not None and po[k].grad.abs().max() > 1e-5:
                assert_rel(p.grad.cpu().numpy(), po[k].grad.numpy(), WHOLE_NET_GRAD_TOL, k)


def test_cross_entropy_kernel(golden_dir):
    g = np.load(os.path.join(golden_dir, "ce_loss.npz"))
    lg = ((golden_input((2, 20, 16, 24), 8) - 0.5) * 6).cuda().requires_grad_(True)
    lb = golden_labels((2, 16, 24), 8).cuda()
    l = CrossEntropyLoss2d()(lg, lb)
    l.backward()
    np.testing.assert_allclose(float(l), float(g["loss"]), rtol=1e-5)
    assert_rel(lg.grad.cpu().numpy(), g["dlogits"], 1e-5)
    lw = CrossEntropyLoss2d(weight=torch.linspace(0.5, 1.5, 20))(lg.detach(), lb)
    np.testing.assert_allclose(float(lw), float(g["loss_weighted"]), rtol=1e-5)
    # all pixels ignored -> nan like torch
    assert torch.isnan(CrossEntropyLoss2d()(lg.detach(), torch.full_like(lb, -1)))


def test_iou_eval_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "iou_kat.npz"))
    for name, n in (("c20", 20), ("c20_big", 20), ("c5_noignore", 5)):
        p8, t8 = torch.from_numpy(g[name + "_pred"]).cuda(), torch.from_numpy(g[name + "_gt"]).cuda()
        for cast in (lambda t: t.long(), lambda t: t):  # int64 (reference API) and the fused u8 path
            ev = iouEval(n)
            ev.addBatch(cast(p8), cast(t8))
            ev.addBatch(cast(t8.flip(0)), cast(t8))
            mean, per = ev.getIoU()
            assert np.array_equal(ev.tp.numpy(), g[name + "_tp"]) and np.array_equal(ev.fp.numpy(), g[name + "_fp"])
            assert np.array_equal(ev.fn.numpy(), g[name + "_fn"])
            np.testing.assert_allclose(per.numpy(), g[name + "_iou"], rtol=1e-12)
    # full config-5 size: 8 x 1024 x 2048 labels, checksum properties + oracle on a slice
    gen = torch.Generator(device="cuda").manual_seed(1)
    p = torch.randint(0, 20, (8, 1, 1024, 2048), device="cuda", generator=gen, dtype=torch.int64)
    t = torch.randint(0, 20, (8, 1, 1024, 2048), device="cuda", generator=gen, dtype=torch.int64)
    ev = iouEval(20)
    ev.addBatch(p, t)
    ev.getIoU()
    cm = ev.confusion_matrix().cpu().numpy()
    assert cm.sum() == p.numel()
    assert np.array_equal(cm.sum(1), np.bincount(t.cpu().numpy().ravel(), minlength=20))
    tp, fp, fn = O.iou_counts(p[:1].cpu().numpy(), t[:1].cpu().numpy(), 20)
    ev1 = iouEval(20)
    ev1.addBatch(p[:1], t[:1])
    ev1.getIoU()
    assert np.array_equal(ev1.tp.numpy(), tp) and np.array_equal(ev1.fp.numpy(), fp) and np.array_equal(ev1.fn.numpy(), fn)
    # empty batch is a no-op
    ev1.addBatch(p[:0], t[:0])


def test_argmax_and_adam():
    lg = (golden_input((2, 20, 16, 24), 9) - 0.5).cuda()
    assert torch.equal(ops.argmax(lg, 20, True), lg.argmax(1))
    n = 10007
    gen = torch.Generator().manual_seed(4)
    p, gr = torch.randn(n, generator=gen), torch.randn(n, generator=gen)
    m, v = torch.zeros(n), torch.zeros(n)
    pc, mc, vc = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
    for step in (1, 2, 3):
        O.adam_step(p, gr, m, v, step)
        ops.adam_step(pc, gr.cuda(), mc, vc, 1e-4, 0.9, 0.999, 1e-4, 1e-6, step)
    np.testing.assert_allclose(pc.cpu().numpy(), p.numpy(), rtol=1e-6, atol=1e-7)


def test_state_dict_roundtrip_and_module_prefix():
    m, sd = _model("fp32")
    out = m.state_dict()
    assert list(out.keys()) == list(sd.keys())
    wrapped = {"module." + k: v for k, v in out.items()}  # DataParallel-saved checkpoint (eval/eval_iou.py:54-65)
    own = m.state_dict()
    for name, param in wrapped.items():
        own[name.split("module.")[-1]].copy_(param)


def test_cross_entropy_nhwc_bf16_fast_path():
    """The vectorised NHWC bf16 CE kernel (logits padded 20 -> 24 channels) vs the oracle on the same bf16 values."""
    gen = torch.Generator().manual_seed(5)
    lg = ((torch.rand(2, 33, 47, 24, generator=gen) - 0.5) * 8).bfloat16()
    lb = golden_labels((2, 33, 47), 9)
    wv = torch.linspace(0.5, 1.5, 20)
    lg_ref = lg[..., :20].float().permute(0, 3, 1, 2).requires_grad_(True)
    lo = O.cross_entropy_2d(lg_ref, lb, weight=wv)
    lo.backward()
    lsum, wsum, dlog = ops.cross_entropy(lg.cuda(), lb.cuda(), wv.cuda(), -1, 20, nchw=False, want_grad=True)
    np.testing.assert_allclose(float(lsum / wsum), float(lo), rtol=2e-4)
    assert_rel(dlog[..., :20].float().cpu().permute(0, 3, 1, 2).numpy(), lg_ref.grad.numpy(), 1e-2)
    assert float(dlog[..., 20:].float().abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_stem_kernels(dtype):
    """erf_stem_fwd / erf_stem_wgrad (DownsamplerBlock(3,16), eval/erfnet.py:67) vs torch on the same stored operands."""
    gen = torch.Generator().manual_seed(17)
    n, h, w = 2, 24, 1024           # full-width rows: the wgrad kernel stages 3 input rows + 1 gradient row in shared memory
    x = torch.zeros(n, h, w, 4)
    x[..., :3] = torch.rand(n, h, w, 3, generator=gen) - 0.3
    x = x.to(dtype)
    wt = torch.randn(13, 3, 3, 3, generator=gen) * 0.3
    bias = torch.randn(13, generator=gen) * 0.1
    xr = x[..., :3].float().permute(0, 3, 1, 2).requires_grad_(False)
    wr = wt.clone().requires_grad_(True)
    br = bias.clone().requires_grad_(True)
    ref = torch.cat([torch.nn.functional.conv2d(xr, wr, br, stride=2, padding=1), torch.nn.functional.max_pool2d(xr, 2, 2)], 1)
    tol = 1e-5 if dtype == torch.float32 else 1e-2
    stats = torch.zeros(32, device="cuda", dtype=torch.float64)
    y = ops.stem_fwd(x.cuda(), wt.cuda(), bias.cuda(), stats=stats)
    assert_rel(y.float().cpu().numpy(), ref.detach().permute(0, 2, 3, 1).numpy(), tol, "stem fwd")
    ys = y.double()
    np.testing.assert_allclose(stats[:16].cpu().numpy(), ys.sum((0, 1, 2)).cpu().numpy(), rtol=1e-5, atol=1e-3)
    np.testing.assert_allclose(stats[16:].cpu().numpy(), (ys * ys).sum((0, 1, 2)).cpu().numpy(), rtol=1e-5, atol=1e-3)
    sc, sh = torch.rand(16, generator=gen) + 0.5, torch.randn(16, generator=gen) * 0.1
    y2 = ops.stem_fwd(x.cuda(), wt.cuda(), bias.cuda(), sc.cuda(), sh.cuda(), relu=True)
    assert_rel(y2.float().cpu().numpy(), (ref.detach() * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).relu().permute(0, 2, 3, 1).numpy(), tol, "stem affine")
    dy = (torch.randn(n, h // 2, w // 2, 16, generator=gen) * 0.5).to(dtype)
    ref.backward(dy.float().permute(0, 3, 1, 2))
    dw0 = torch.full((13, 3, 3, 3), 0.25, device="cuda")            # accumulate semantics
    dw, db = ops.stem_wgrad(x.cuda(), dy.cuda(), dw_out=dw0)
    assert_rel((dw - 0.25).cpu().numpy(), wr.grad.numpy(), 1e-4 if dtype == torch.float32 else 2e-3, "stem dw")
    assert_rel(db.cpu().numpy(), br.grad.numpy(), 1e-4 if dtype == torch.float32 else 2e-3, "stem db")


def test_full_resolution_forward_and_iou_config5():
    """BASELINE config 5 shape (1024 x 2048, here one frame): eval forward on the tcgen05 path vs the fp32 oracle, fused argmax, and the
    confusion histogram of those predictions vs the oracle's integer counts."""
    from erfnet_pytorch_b200.iouEval import iouEval
    m = E.ERFNet(20)
    sd = O.make_state_dict(O.erfnet_param_shapes(20), 6)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x = golden_input((1, 3, 1024, 2048), 31)
    labels = golden_labels((1, 1024, 2048), 32, n_classes=20, ignore=19)
    with torch.no_grad():
        ref = O.erfnet_forward(x, sd)
        y = m(x.cuda())
    assert_rel(y.cpu().numpy(), ref.numpy(), 2e-2, "1024x2048 logits vs oracle")
    pred = ops.argmax(y.contiguous(), 20, True)                      # [1,H,W] int64, first maximum wins like torch.max(1)
    assert torch.equal(pred.cpu(), y.cpu().argmax(1))
    # north_star gate: >= 99.9 % argmax agreement.  With RANDOM weights the top-2 margin of many pixels is below the bf16 error, so the
    # gate is asserted in fp32 check mode, and in bf16 for every pixel whose oracle margin exceeds twice the measured logit error.
    agree = (pred.cpu() == ref.argmax(1)).float().mean().item()
    top2 = ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1])
    err = (y.cpu() - ref).abs().max().item()
    decided = margin > 2 * err
    assert bool((pred.cpu() == ref.argmax(1))[decided].all()) and decided.float().mean().item() > 0.5
    assert agree >= 0.98, f"bf16 argmax agreement with the fp32 oracle {agree:.5f}"
    m.precision = "fp32"
    with torch.no_grad():
        y32 = m(x.cuda())
    assert_rel(y32.cpu().numpy(), ref.numpy(), 1e-4, "1024x2048 fp32 check mode vs oracle")
    agree32 = (y32.cpu().argmax(1) == ref.argmax(1)).float().mean().item()
    assert agree32 >= 0.999, f"fp32 argmax agreement {agree32:.5f}"
    ev = iouEval(20)
    ev.addBatch(pred.unsqueeze(1), labels.unsqueeze(1).cuda())
    tp, fp, fn = O.iou_counts(pred.cpu().numpy(), labels.numpy(), 20, 19)
    np.testing.assert_array_equal(ev.tp.cpu().numpy().astype(np.int64), tp)
    np.testing.assert_array_equal(ev.fp.cpu().numpy().astype(np.int64), fp)
    np.testing.assert_array_equal(ev.fn.cpu().numpy().astype(np.int64), fn)


@pytest.mark.parametrize("shape", [(2, 3, 64, 128), (1, 3, 8, 8), (1, 3, 40, 24)])
def test_stem_from_nchw_matches_the_layout_pass(shape):
    """Inference: the stem reads the NCHW fp32 image itself.  In fp32 check mode the result is bit-identical to layout pass + NHWC stem
    (same arithmetic order); in bf16 it skips the bf16 rounding of the input pixels, so it is compared within the bf16 gate."""
    m = E.ERFNet(20)
    m.load_state_dict(O.make_state_dict(O.erfnet_param_shapes(20), 4))
    m = m.cuda().eval()
    x = torch.rand(shape, generator=torch.Generator().manual_seed(13)).cuda()
    for precision in ("fp32", "bf16"):
        m.precision = precision
        ctx = E._Ctx(precision, False)
        with torch.no_grad():
            a = m.encoder.initial_block._fwd_eval_nchw(x, ctx)
            b = m.encoder.initial_block._fwd_eval(E.input_to_nhwc(x, 4, ctx.dtype), ctx)
        assert a.shape == b.shape == (shape[0], shape[2] // 2, shape[3] // 2, 16) and a.dtype == b.dtype
        if precision == "fp32":
            assert torch.equal(a, b)
        else:
            assert (a.float() - b.float()).abs().max().item() <= 2e-2 * b.float().abs().max().item()


@pytest.mark.parametrize("shape", [(2, 3, 8, 12), (1, 3, 6, 10), (6, 3, 64, 128)])
def test_image_ingest_nchw3_to_nhwc4_matches_generic(shape):
    """The vectorised 3-plane fp32 -> NHWC bf16 (4-channel pixels) conversion (four pixels per thread; HW % 4 == 0) against the definition,
    and the generic kernel for sizes it does not take (HW % 4 != 0)."""
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(7)).cuda()
    got = ops.nchw_to_nhwc(x, 4, torch.bfloat16)
    ref = torch.zeros(shape[0], shape[2], shape[3], 4, device="cuda", dtype=torch.bfloat16)
    ref[..., :3] = x.permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(got, ref)
