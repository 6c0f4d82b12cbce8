"""GPU: the ImageNet net and training step (BASELINE config 3), row-wise CE / top-k, flat SGD, the mean-teacher update and the I/O-side
kernels (uint8 ingest, label tables, colouring) through the C ABI."""
import pytest
import torch

from tests.imagenet_cases import build_imagenet, check_ce_topk, check_imagenet_flat_trainer, check_imagenet_net, check_pretrained_encoder_handoff
from tests.io_cases import check_ingest, check_iou_one_hot, check_labels, check_mean_teacher, check_model_accepts_uint8
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu


def test_imagenet_net_fp32_golden(golden_dir):
    check_imagenet_net(golden_dir, "cuda")


def test_imagenet_flat_trainer(golden_dir):
    check_imagenet_flat_trainer(golden_dir, "cuda")


def test_pretrained_encoder_handoff():
    check_pretrained_encoder_handoff("cuda")


def test_ce_rows_topk(golden_dir):
    check_ce_topk(golden_dir, "cuda")


def test_imagenet_bf16_vs_fp32():
    """Batch 8 at 224 x 224, eval mode: bf16 storage vs the fp32 check mode of the same kernels."""
    x = torch.randn(8, 3, 224, 224, generator=torch.Generator().manual_seed(9)).cuda()
    outs = {}
    for precision in ("fp32", "bf16"):
        m = build_imagenet("cuda", precision).eval()
        with torch.no_grad():
            outs[precision] = m(x).float().cpu().numpy()
    assert outs["bf16"].shape == (8, 1000)
    assert_rel(outs["bf16"], outs["fp32"], 3e-2, "imagenet bf16 vs fp32")


def test_ingest_labels_colorize():
    check_ingest("cuda")
    check_labels("cuda")
    check_model_accepts_uint8("cuda", "fp32")
    check_model_accepts_uint8("cuda", "bf16")


def test_iou_one_hot_inputs():
    check_iou_one_hot("cuda")


def test_mean_teacher():
    check_mean_teacher("cuda")


def test_mean_teacher_training_vs_reference(golden_dir):
    from tests.mean_teacher_cases import check_mean_teacher_training
    check_mean_teacher_training(golden_dir, "cuda", "fp32")


def test_tail_kernels_against_torch():
    """Kernel-level parity of the ImageNet-tail / optimiser kernels on ragged sizes: AvgPool2d(k) with a multi-pixel output, Linear with sizes
    that are not tile multiples, SGD with momentum + weight decay over an odd-length buffer, EMA."""
    import numpy as np
    import torch.nn.functional as F
    from erfnet_pytorch_b200 import ops
    g = torch.Generator().manual_seed(21)
    # AvgPool2d(5, 1, 0) on 9 x 11, 24 channels (pitch 24), fp32 and bf16
    x = torch.randn(2, 9, 11, 24, generator=g)
    ref = F.avg_pool2d(x.permute(0, 3, 1, 2), 5, 1, 0).permute(0, 2, 3, 1)
    for dtype, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        got = ops.avgpool_fwd(x.cuda().to(dtype), 24, 5)
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=tol, atol=tol)
    dout = torch.randn(2, 5, 7, 24, generator=g)
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.avg_pool2d(xr, 5, 1, 0).backward(dout.permute(0, 3, 1, 2))
    dx = ops.avgpool_bwd(dout.cuda(), 9, 11, 24, 5, torch.float32)
    np.testing.assert_allclose(dx.cpu().numpy(), xr.grad.permute(0, 2, 3, 1).numpy(), rtol=1e-5, atol=1e-6)
    # Linear 19 -> 37 on 5 rows
    xl, w, b = torch.randn(5, 19, generator=g), torch.randn(37, 19, generator=g), torch.randn(37, generator=g)
    y = ops.linear_fwd(xl.cuda(), w.cuda(), b.cuda())
    np.testing.assert_allclose(y.cpu().numpy(), F.linear(xl, w, b).numpy(), rtol=1e-5, atol=1e-5)
    dy = torch.randn(5, 37, generator=g)
    dw0, db0 = torch.ones(37, 19).cuda(), torch.ones(37).cuda()
    dxl, dw, db = ops.linear_bwd(xl.cuda(), w.cuda(), dy.cuda(), True, dw0, db0)       # accumulates into dw0 / db0
    np.testing.assert_allclose(dxl.cpu().numpy(), (dy @ w).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(dw.cpu().numpy(), (dy.t() @ xl + 1).numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(db.cpu().numpy(), (dy.sum(0) + 1).numpy(), rtol=1e-5, atol=1e-5)
    # SGD(0.1, momentum 0.9, wd 1e-2) on 10007 values, 3 steps, vs torch.optim.SGD
    p = torch.randn(10007, generator=g)
    pr = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([pr], 0.1, momentum=0.9, weight_decay=1e-2)
    pc, buf = p.clone().cuda(), torch.zeros(10007).cuda()
    for step in range(3):
        gr = torch.randn(10007, generator=g)
        pr.grad = gr.clone()
        opt.step()
        ops.sgd_step(pc, gr.cuda(), buf, 0.1, 0.9, 1e-2, step == 0)
    np.testing.assert_allclose(pc.cpu().numpy(), pr.detach().numpy(), rtol=1e-5, atol=1e-6)
    # EMA
    t, s2 = torch.randn(1001, generator=g), torch.randn(1001, generator=g)
    tc = t.clone().cuda()
    ops.ema_update(tc, s2.cuda(), 0.01)
    np.testing.assert_allclose(tc.cpu().numpy(), (t * 0.99 + 0.01 * s2).numpy(), rtol=1e-6, atol=1e-7)
