"""GPU: the single-launch train-mode BatchNorm kernels (csrc/bn_fused.cu) against a float64 restatement of nn.BatchNorm2d's
training arithmetic (eval/erfnet.py:19,37,55,105: BatchNorm2d(C, eps=1e-3) -> [Dropout2d] -> [+ input] -> ReLU and its autograd) and against
the round-1 two-launch kernels they replace."""
import numpy as np
import pytest
import torch

from erfnet_pytorch_b200 import ops

pytestmark = pytest.mark.gpu

EPS, MOM = 1e-3, 0.1


def _rand(shape, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * 1.5 + 0.3).to(dtype).cuda()


def _ref_forward(x, gamma, beta, drop, res, relu):
    xd = x.double()
    n, h, w, c = x.shape
    mean = xd.mean((0, 1, 2))
    var = xd.var((0, 1, 2), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + EPS)
    y = (xd - mean) * invstd * gamma.double() + beta.double()
    if drop is not None:
        y = y * drop.double().view(n, 1, 1, c)
    if res is not None:
        y = y + res.double()
    if relu:
        y = y.relu()
    cnt = n * h * w
    return y, mean, invstd, var * cnt / (cnt - 1)


# sizes: the bench tensors' channel counts at small extents, a ragged pixel count, and one large enough that a thread owns more items
# than the backward kernel can keep on chip (the re-read path)
CASES = [(2, 8, 16, 128), (1, 7, 9, 64), (3, 16, 32, 16), (2, 64, 128, 64), (6, 128, 256, 64), (5, 40, 100, 64), (3, 50, 60, 16)]


@pytest.mark.parametrize("shape", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_drop,with_res", [(False, False), (True, True)])
def test_bn_forward_single_launch(shape, dtype, with_drop, with_res):
    n, h, w, c = shape
    x = _rand(shape, 1, dtype)
    gamma = (torch.rand(c, generator=torch.Generator().manual_seed(2)) + 0.5).cuda()
    beta = (torch.rand(c, generator=torch.Generator().manual_seed(3)) - 0.5).cuda()
    drop = ((torch.rand(n, c, generator=torch.Generator().manual_seed(4)) > 0.3).float() / 0.7).cuda() if with_drop else None
    res = _rand(shape, 5, dtype) if with_res else None
    outs = []
    for fused in (True, False):
        ops.BN_FUSED = fused
        try:
            stats = ops.bn_stats_buf(c, x.device)
            ops.channel_stats(x, stats)
            rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
            before = ops.LAUNCHES
            out, mean, invstd = ops.bn_apply(x, stats, n * h * w, gamma, beta, rm, rv, MOM, EPS, drop=drop, res=res, relu=True)
            launches = ops.LAUNCHES - before
            outs.append((out, mean, invstd, rm, rv))
            assert launches == (1 if fused else 2)
        finally:
            ops.BN_FUSED = True
    y, mean, invstd, unb = _ref_forward(x, gamma, beta, drop, res, True)
    out, m, i, rm, rv = outs[0]
    tol = 2e-5 if dtype == torch.float32 else 1.2e-2
    scale = float(y.abs().max())
    assert float((out.double() - y).abs().max()) <= tol * scale
    np.testing.assert_allclose(m.cpu().numpy(), mean.cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(i.cpu().numpy(), invstd.cpu().numpy(), rtol=1e-5)
    np.testing.assert_allclose(rm.cpu().numpy(), (MOM * mean).cpu().numpy(), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(rv.cpu().numpy(), (1 - MOM + MOM * unb).cpu().numpy(), rtol=1e-5)
    # same arithmetic as finalize + apply: identical results
    for a, b in zip(outs[0], outs[1]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("shape", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_drop,want_masked", [(False, False), (True, True), (False, True)])
def test_bn_backward_single_launch(shape, dtype, with_drop, want_masked):
    n, h, w, c = shape
    y = _rand(shape, 11, dtype)                       # pre-BN activation
    g = _rand(shape, 12, dtype) * 0.1                 # upstream gradient
    act = _rand(shape, 13, dtype)                     # stored post-activation tensor: mask = act > 0
    gamma = (torch.rand(c, generator=torch.Generator().manual_seed(14)) + 0.5).cuda()
    drop = ((torch.rand(n, c, generator=torch.Generator().manual_seed(15)) > 0.3).float() / 0.7).cuda() if with_drop else None
    yd = y.double()
    mean = yd.mean((0, 1, 2))
    invstd = 1.0 / torch.sqrt(yd.var((0, 1, 2), unbiased=False) + EPS)
    meanf, invf = mean.float(), invstd.float()
    # float64 reference of BatchNorm backward with upstream gb = g * (act > 0) * drop
    gm = g.double() * (act.double() > 0)
    gb = gm * drop.double().view(n, 1, 1, c) if drop is not None else gm
    xhat = (yd - meanf.double()) * invf.double()
    m = n * h * w
    s1, s2 = gb.sum((0, 1, 2)), (gb * xhat).sum((0, 1, 2))
    dy_ref = gamma.double() * invf.double() * (gb - s1 / m - xhat * s2 / m)
    res = []
    for fused in (True, False):
        ops.BN_FUSED = fused
        try:
            before = ops.LAUNCHES
            r = ops.bn_bwd(g, y, gamma, meanf, invf, mask=act, drop=drop, want_masked=want_masked)
            res.append(r)
            assert ops.LAUNCHES - before == (1 if fused else 2)
        finally:
            ops.BN_FUSED = True
    dy, dgamma, dbeta = res[0][:3]
    tol = 5e-5 if dtype == torch.float32 else 1.2e-2
    assert float((dy.double() - dy_ref).abs().max()) <= tol * float(dy_ref.abs().max())
    np.testing.assert_allclose(dgamma.cpu().numpy(), s2.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(s2.abs().max()))
    np.testing.assert_allclose(dbeta.cpu().numpy(), s1.cpu().numpy(), rtol=2e-4, atol=2e-4 * float(s1.abs().max()))
    if want_masked:
        assert torch.equal(res[0][3], gm.to(dtype))                     # exact: a masked copy
        assert torch.equal(res[0][3], res[1][3])
    # against the two-launch kernels: same formulas, the reductions only differ in summation order
    assert float((res[0][0].double() - res[1][0].double()).abs().max()) <= (1e-5 if dtype == torch.float32 else 1e-2) * float(dy_ref.abs().max())


def test_bn_backward_repeated_launches_reuse_nothing_stale():
    """Every call takes a fresh (zeroed) sums + barrier slot; 50 back-to-back launches on one stream must all agree (a stale counter
    would let a CTA pass the barrier early and read partial sums)."""
    shape = (6, 64, 128, 128)
    n, h, w, c = shape
    y, g, act = _rand(shape, 21, torch.bfloat16), _rand(shape, 22, torch.bfloat16), _rand(shape, 23, torch.bfloat16)
    gamma = torch.ones(c, device="cuda")
    yd = y.double()
    meanf = yd.mean((0, 1, 2)).float()
    invf = (1.0 / torch.sqrt(yd.var((0, 1, 2), unbiased=False) + EPS)).float()
    arena = ops.ScratchArena(y.device)
    ops.ARENA = arena
    try:
        arena.reset()
        first = ops.bn_bwd(g, y, gamma, meanf, invf, mask=act)
        for i in range(50):
            if arena.cursor + 2 * c + 1 > arena.buf.numel():
                arena.reset()
            r = ops.bn_bwd(g, y, gamma, meanf, invf, mask=act)
            assert float((r[0].float() - first[0].float()).abs().max()) <= 2e-2 * float(first[0].float().abs().max()), i
            np.testing.assert_allclose(r[1].cpu().numpy(), first[1].cpu().numpy(), rtol=1e-3, atol=1e-2)
    finally:
        ops.ARENA = None


@pytest.mark.parametrize("shape", [(2, 8, 16, 128), (6, 128, 256, 64), (3, 16, 32, 16)])
def test_bn_backward_mask_recomputed_from_y(shape):
    """For a plain BN -> ReLU the bf16 kernel derives the ReLU mask from y with the forward pass's own expression instead of reading the
    activation: g * (mask > 0) must be IDENTICAL to the one taken from the stored activation."""
    n, h, w, c = shape
    y = _rand(shape, 31, torch.bfloat16)
    g = _rand(shape, 32, torch.bfloat16) * 0.1
    gamma = (torch.rand(c, generator=torch.Generator().manual_seed(33)) + 0.5).cuda()
    gamma[::3] *= -1.0                                                   # negative scales flip the inequality: must be handled
    beta = (torch.rand(c, generator=torch.Generator().manual_seed(34)) - 0.5).cuda()
    stats = ops.bn_stats_buf(c, y.device)
    ops.channel_stats(y, stats)
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    act, mean, invstd = ops.bn_apply(y, stats, n * h * w, gamma, beta, rm, rv, MOM, EPS, relu=True)
    a = ops.bn_bwd(g, y, gamma, mean, invstd, mask=act, want_masked=True)
    b = ops.bn_bwd(g, y, gamma, mean, invstd, mask=act, want_masked=True, beta_for_mask=beta)
    assert 0.2 < float((act > 0).float().mean()) < 0.8
    assert torch.equal(a[3], b[3])
    assert float((a[0].float() - b[0].float()).abs().max()) <= 1e-2 * float(a[0].float().abs().max())
    np.testing.assert_allclose(a[1].cpu().numpy(), b[1].cpu().numpy(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(a[2].cpu().numpy(), b[2].cpu().numpy(), rtol=1e-4, atol=1e-3)
