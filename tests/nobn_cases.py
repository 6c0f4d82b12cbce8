"""Shared body of the BN-free (fork default, train/erfnet_blocks.py) block tests: train-mode forward + backward of
non_bottleneck_1d / DownsamplerBlock / UpsamplerBlock with use_batchnorm=False against tests/golden/blocks_nobn.npz
(outputs of the reference's own fork blocks).  CPU: kernels emulated; GPU: real kernels."""
import os

import numpy as np
import torch

from erfnet_pytorch_b200 import erfnet as E
from erfnet_pytorch_b200 import erfnet_blocks as EB
from oracle import erfnet_oracle as O
from oracle.make_golden import NOBN_DOWN, NOBN_NB, NOBN_UP, golden_input
from tests.test_host_logic import assert_rel as _assert_rel_max

assert_rel = _assert_rel_max


def _load(blk, seed, device, precision):
    blk.precision = precision
    bsd = O.make_state_dict({"blk." + k: tuple(v.shape) for k, v in blk.state_dict().items()}, seed=seed)
    blk.load_state_dict({k[4:]: v for k, v in bsd.items()})
    blk.to(device).train()
    return blk


def assert_rel_l2(actual, desired, tol, msg=""):
    """||a-d||_2 <= tol * ||d||_2: the metric for bf16 gradients, where a handful of flipped ReLU masks make single-pixel
    errors of order max|ref| unavoidable."""
    a, d = np.asarray(actual, np.float64), np.asarray(desired, np.float64)
    err, ref = np.linalg.norm(a - d), np.linalg.norm(d) + 1e-30
    assert err <= tol * ref, f"{msg}: ||diff||={err:.3e} vs tol*||ref||={tol * ref:.3e}"


def check_blocks_nobn(golden_dir, device, precision="fp32", tol=1e-3, names=None, l2=False):
    global assert_rel
    assert_rel = assert_rel_l2 if l2 else _assert_rel_max
    g = np.load(os.path.join(golden_dir, "blocks_nobn.npz"))
    for name, (c, d, hw, p, use_dropout) in NOBN_NB.items():
        if names is not None and name not in names:
            continue
        blk = _load(E.non_bottleneck_1d(c, p, d, use_dropout=use_dropout, use_batchnorm=False), 13, device, precision)
        x = (golden_input((2, c) + hw, 40 + d) - 0.3).to(device).requires_grad_(True)
        rm0 = blk.bn1.running_mean.clone()
        fixed = {blk: torch.from_numpy(g[name + "_drop"])} if name + "_drop" in g.files else {}
        with E.fixed_dropout(fixed):
            y = blk(x)
        y.backward((golden_input(tuple(y.shape), 87) - 0.5).to(device))
        assert_rel(y.detach().cpu().numpy(), g[name + "_y"], tol / 4, name + " y")
        assert_rel(x.grad.cpu().numpy(), g[name + "_dx"], tol, name + " dx")
        for cn in ("conv3x1_1", "conv1x3_1", "conv3x1_2", "conv1x3_2"):
            m = getattr(blk, cn)
            assert_rel(m.bias.grad.cpu().numpy(), g[f"{name}_db_{cn}"], tol, f"{name} db {cn}")
            if f"{name}_dw_{cn}" in g.files:
                assert_rel(m.weight.grad.cpu().numpy(), g[f"{name}_dw_{cn}"], tol, f"{name} dw {cn}")
        assert blk.bn1.weight.grad is None and blk.bn2.bias.grad is None           # unused, as in the reference
        assert torch.equal(blk.bn1.running_mean, rm0)
    for table, cls, seed, gseed in ((NOBN_DOWN, E.DownsamplerBlock, 14, 88), (NOBN_UP, E.UpsamplerBlock, 15, 89)):
        for name, (ci, co, hw) in table.items():
            if names is not None and name not in names:
                continue
            blk = _load(cls(ci, co, use_batchnorm=False), seed, device, precision)
            x = (golden_input((2, ci) + hw, 50 + ci) - 0.3).to(device).requires_grad_(True)
            y = blk(x)
            y.backward((golden_input(tuple(y.shape), gseed) - 0.5).to(device))
            assert_rel(y.detach().cpu().numpy(), g[name + "_y"], tol / 4, name + " y")
            assert_rel(x.grad.cpu().numpy(), g[name + "_dx"], tol, name + " dx")
            assert_rel(blk.conv.weight.grad.cpu().numpy(), g[name + "_dw"], tol, name + " dw")
            assert_rel(blk.conv.bias.grad.cpu().numpy(), g[name + "_db"], tol, name + " db")
            # eval mode of the BN-free block = the same arithmetic without autograd
            blk.eval()
            with torch.no_grad():
                assert_rel(blk(x.detach()).cpu().numpy(), g[name + "_y"], tol / 4, name + " eval")
    for name, cls in (("decblock64_16", EB.DecoderBlock), ("decblocksmall64_16", EB.DecoderBlockSmall)):
        if names is not None and name not in names:
            continue
        blk = _load(cls(64, 16), 16, device, precision)
        x = (golden_input((2, 64, 6, 8), 61) - 0.3).to(device).requires_grad_(True)
        y = blk(x)
        y.backward((golden_input(tuple(y.shape), 90) - 0.5).to(device))
        assert_rel(y.detach().cpu().numpy(), g[name + "_y"], tol / 4, name + " y")
        assert_rel(x.grad.cpu().numpy(), g[name + "_dx"], tol, name + " dx")
        assert_rel(blk.layers[0].conv.weight.grad.cpu().numpy(), g[name + "_dw_up"], tol, name + " dw_up")
        assert_rel(blk.layers[-1].conv1x3_2.bias.grad.cpu().numpy(), g[name + "_db_last"], tol, name + " db_last")


def check_erfnet_nobn(golden_dir, device, precision="fp32", tol=1e-4):
    """eval/erfnet_nobn.ERFNet(20): checkpoint layout and eval logits (full net and only_encode) vs the reference's own module."""
    import os
    from erfnet_pytorch_b200 import erfnet_nobn
    from oracle.make_golden import golden_input
    g = np.load(os.path.join(golden_dir, "erfnet_nobn.npz"))
    m = erfnet_nobn.ERFNet(20)
    assert [f"{k} {tuple(v.shape)}" for k, v in m.state_dict().items()] == list(g["keys"])
    m.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=71), strict=True)
    m.precision = precision
    m = m.to(device).eval()
    x = golden_input((1, 3, 64, 128), 141).to(device)
    with torch.no_grad():
        _assert_rel_max(m(x).cpu().numpy(), g["logits"], tol, "erfnet_nobn logits")
        _assert_rel_max(m(x, only_encode=True).cpu().numpy(), g["enc_logits"], tol, "erfnet_nobn encoder logits")
