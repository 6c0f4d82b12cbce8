"""CPU: pin oracle/erfnet_oracle.py against the golden outputs of the reference (tests/golden, made by
oracle/make_golden.py by executing /root/reference).  fp32 vs fp32 on the same torch build -> tight tolerances."""
import os

import numpy as np
import torch

from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input, golden_labels

TOL = dict(rtol=1e-4, atol=2e-5)


def _blk_sd(shapes, seed):
    return O.make_state_dict({"blk." + k: s for k, s in shapes.items()}, seed=seed)


def _nb_shapes(c):
    full = O.erfnet_param_shapes(20)
    pref = {16: "decoder.layers.4.", 64: "encoder.layers.1.", 128: "encoder.layers.7."}[c]
    return {k[len(pref):]: s for k, s in full.items() if k.startswith(pref)}


def test_state_dict_layout(golden_dir):
    lines = open(os.path.join(golden_dir, "state_dict_keys.txt")).read().strip().split("\n")
    shapes = O.erfnet_param_shapes(20)
    assert len(lines) == 345 == len(shapes)
    for line, (k, s) in zip(lines, shapes.items()):
        name, rest = line.split(" ", 1)
        assert name == k and rest.startswith(str(tuple(s)))


def test_whole_net_eval(golden_dir):
    g = np.load(os.path.join(golden_dir, "erfnet_eval_64x128.npz"))
    sd = O.make_state_dict(O.erfnet_param_shapes(20), 0)
    x = golden_input((1, 3, 64, 128), 0)
    with torch.no_grad():
        np.testing.assert_allclose(O.erfnet_forward(x, sd).numpy(), g["logits"], **TOL)
        np.testing.assert_allclose(O.erfnet_forward(x, sd, only_encode=True).numpy(), g["enc_logits"], **TOL)


def test_blocks(golden_dir):
    g = np.load(os.path.join(golden_dir, "blocks.npz"))
    cases = {"nb64_d1": (64, 1, (8, 24)), "nb128_d2": (128, 2, (8, 12)), "nb128_d16": (128, 16, (18, 20)),
             "nb16_d1": (16, 1, (16, 24)), "nb128_d4": (128, 4, (8, 16)), "nb128_d8": (128, 8, (10, 18))}
    for name, (c, d, hw) in cases.items():
        sd = _blk_sd(_nb_shapes(c), 3)
        x = golden_input((2, c) + hw, 10 + d) - 0.3
        with torch.no_grad():
            np.testing.assert_allclose(O.non_bottleneck_1d(x, sd, "blk", d).numpy(), g[name + "_eval"], **TOL)
        for k in sd:
            if sd[k].is_floating_point() and "running" not in k:
                sd[k].requires_grad_(True)
        x.requires_grad_(True)
        y = O.non_bottleneck_1d(x, sd, "blk", d, training=True)
        y.backward(golden_input(tuple(y.shape), 77) - 0.5)
        np.testing.assert_allclose(y.detach().numpy(), g[name + "_train"], **TOL)
        np.testing.assert_allclose(x.grad.numpy(), g[name + "_train_dx"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(sd["blk.conv1x3_2.weight"].grad.numpy(), g[name + "_train_dw_conv1x3_2"], rtol=1e-3, atol=1e-3)
        np.testing.assert_allclose(sd["blk.bn1.weight"].grad.numpy(), g[name + "_train_dgamma1"], rtol=1e-3, atol=1e-3)
        np.testing.assert_allclose(sd["blk.bn2.running_mean"].numpy(), g[name + "_train_rm2"], **TOL)
        np.testing.assert_allclose(sd["blk.bn2.running_var"].numpy(), g[name + "_train_rv2"], **TOL)
    full = O.erfnet_param_shapes(20)
    for name, (ci, co, hw, pref) in {"down3_16": (3, 16, (16, 32), "encoder.initial_block."),
                                      "down16_64": (16, 64, (8, 16), "encoder.layers.0."),
                                      "down64_128": (64, 128, (8, 8), "encoder.layers.6.")}.items():
        sd = _blk_sd({k[len(pref):]: s for k, s in full.items() if k.startswith(pref)}, 4)
        x = golden_input((2, ci) + hw, 20 + ci) - 0.3
        with torch.no_grad():
            np.testing.assert_allclose(O.downsampler_block(x, sd, "blk").numpy(), g[name + "_eval"], **TOL)
        sd["blk.conv.weight"].requires_grad_(True)
        x.requires_grad_(True)
        y = O.downsampler_block(x, sd, "blk", training=True)
        y.backward(golden_input(tuple(y.shape), 78) - 0.5)
        np.testing.assert_allclose(y.detach().numpy(), g[name + "_train"], **TOL)
        np.testing.assert_allclose(x.grad.numpy(), g[name + "_train_dx"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(sd["blk.conv.weight"].grad.numpy(), g[name + "_train_dw"], rtol=1e-3, atol=1e-3)
    for name, (ci, co, hw, pref) in {"up128_64": (128, 64, (4, 8), "decoder.layers.0."),
                                      "up64_16": (64, 16, (8, 8), "decoder.layers.3.")}.items():
        sd = _blk_sd({k[len(pref):]: s for k, s in full.items() if k.startswith(pref)}, 5)
        x = golden_input((2, ci) + hw, 30 + ci) - 0.3
        with torch.no_grad():
            np.testing.assert_allclose(O.upsampler_block(x, sd, "blk").numpy(), g[name + "_eval"], **TOL)


def test_train_step_losses(golden_dir):
    """Two Adam steps of the full net (train-mode BN, dropout p=0) track the reference."""
    g = np.load(os.path.join(golden_dir, "train_step_32x64.npz"))
    sd = O.make_state_dict(O.erfnet_param_shapes(20), 0)
    params = {k: v for k, v in sd.items() if v.is_floating_point() and "running" not in k}
    for v in params.values():
        v.requires_grad_(True)
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    vv = {k: torch.zeros_like(v) for k, v in params.items()}
    x = golden_input((2, 3, 32, 64), 5)
    y = golden_labels((2, 32, 64), 5)
    losses = []
    for step in (1, 2):
        for v in params.values():
            v.grad = None
        out = O.erfnet_forward(x, sd, training=True, dropprob=0.0)
        loss = O.cross_entropy_2d(out, y)
        loss.backward()
        if step == 1:
            np.testing.assert_allclose(out.detach().numpy(), g["logits_step0"], **TOL)
            for k in g.files:
                if k.startswith("grad0/"):
                    np.testing.assert_allclose(params[k[6:]].grad.numpy(), g[k], rtol=2e-3, atol=2e-5)
            assert sorted(k for k, v in params.items() if v.grad is None) == sorted(g["no_grad_keys"].tolist())
        with torch.no_grad():
            for k, p in params.items():
                if p.grad is not None:
                    O.adam_step(p, p.grad, m[k], vv[k], step)
        losses.append(float(loss))
    np.testing.assert_allclose(losses, g["losses"], rtol=1e-5)
    for k in g.files:
        if k.startswith("param2/"):
            np.testing.assert_allclose(params[k[7:]].detach().numpy(), g[k], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(sd["encoder.layers.9.bn2.running_mean"].numpy(), g["rm/encoder.layers.9.bn2.running_mean"], **TOL)
    np.testing.assert_allclose(sd["decoder.layers.3.bn.running_var"].numpy(), g["rv/decoder.layers.3.bn.running_var"], **TOL)


def test_ce_loss(golden_dir):
    g = np.load(os.path.join(golden_dir, "ce_loss.npz"))
    lg = ((golden_input((2, 20, 16, 24), 8) - 0.5) * 6).requires_grad_(True)
    lb = golden_labels((2, 16, 24), 8)
    l = O.cross_entropy_2d(lg, lb)
    l.backward()
    np.testing.assert_allclose(float(l), float(g["loss"]), rtol=1e-6)
    np.testing.assert_allclose(lg.grad.numpy(), g["dlogits"], rtol=1e-4, atol=1e-8)
    lw = O.cross_entropy_2d(lg.detach(), lb, weight=torch.linspace(0.5, 1.5, 20))
    np.testing.assert_allclose(float(lw), float(g["loss_weighted"]), rtol=1e-6)


def test_iou_kat(golden_dir):
    g = np.load(os.path.join(golden_dir, "iou_kat.npz"))
    for name, n in (("c20", 20), ("c20_big", 20), ("c5_noignore", 5)):
        p, t = g[name + "_pred"].astype(np.int64), g[name + "_gt"].astype(np.int64)
        tp, fp, fn = O.iou_counts(p, t, n)
        tp2, fp2, fn2 = O.iou_counts(t[::-1], t, n)
        assert np.array_equal((tp + tp2).astype(np.float64), g[name + "_tp"])
        assert np.array_equal((fp + fp2).astype(np.float64), g[name + "_fp"])
        assert np.array_equal((fn + fn2).astype(np.float64), g[name + "_fn"])
        mean, per = O.iou_from_counts(tp + tp2, fp + fp2, fn + fn2)
        np.testing.assert_allclose(per, g[name + "_iou"], rtol=1e-12)
        np.testing.assert_allclose(mean, float(g[name + "_mean"]), rtol=1e-12)


def _fork_oracle_cases(d):
    """name -> (oracle loss fn, operands, indices of the operands that receive gradients)"""
    return {
        "l1": (O.l1_loss_masked, (d["out"], d["target"]), (0,)),
        "loglik": (lambda c, v, t: O.loglik_loss_masked(c, v, t, 1e-4), (d["out"][:, 0], d["var"], d["target"]), (0, 1)),
        "l1cp": (O.l1_loss_classprob_masked, (d["prob"], d["cost"], d["target"]), (0, 1)),
        "loglikcp": (lambda p, c, v, t: O.loglik_loss_classprob_masked(p, c, v, t, 1e-4), (d["prob"], d["cost"], d["cvar"], d["target"]), (0, 1, 2)),
        "mse": (lambda a, b: O.mse_loss_weighted(a, b, 3.0, 0.25), (d["out"], d["teacher"]), (0, 1)),
    }


def test_fork_losses(golden_dir):
    """Masked regression losses + accuracy metrics of the fork vs the reference's own classes (fork_losses.npz)."""
    from oracle.make_golden import fork_loss_inputs
    z = np.load(os.path.join(golden_dir, "fork_losses.npz"))
    d = fork_loss_inputs()
    for name, (fn, args, gidx) in _fork_oracle_cases(d).items():
        leaves = [a.clone().requires_grad_(i in gidx) for i, a in enumerate(args)]
        loss = fn(*leaves)
        loss.backward()
        np.testing.assert_allclose(loss.item(), z[name + "_loss"], rtol=1e-5)
        for i in gidx:
            np.testing.assert_allclose(leaves[i].grad.numpy(), z[f"{name}_g{i}"], rtol=1e-4, atol=1e-7, err_msg=f"{name} grad {i}")
    nc, nt, cc, ct = O.accuracy_counts(d["prob"], d["labels"], d["prob"].shape[1])
    np.testing.assert_allclose(nc / nt, z["acc"], rtol=1e-6)
    with np.errstate(invalid="ignore"):
        ca = cc / ct
    np.testing.assert_allclose(ca, z["class_acc"], rtol=1e-6)
    np.testing.assert_allclose(np.nanmean(ca), z["mean_acc"], rtol=1e-6)


def test_blocks_without_batchnorm(golden_dir):
    """Oracle blocks with use_batchnorm=False vs the fork's own BN-free blocks (train/erfnet_blocks.py), train mode."""
    from oracle.make_golden import NOBN_DOWN, NOBN_NB, NOBN_UP
    g = np.load(os.path.join(golden_dir, "blocks_nobn.npz"))

    def leaves(shapes, seed):
        sd = _blk_sd(shapes, seed)
        return {k: (v.clone().requires_grad_(True) if (v.dtype.is_floating_point and "running" not in k) else v) for k, v in sd.items()}

    for name, (c, d, hw, p, use_dropout) in NOBN_NB.items():
        sd = leaves(_nb_shapes(c), 13)
        x = (golden_input((2, c) + hw, 40 + d) - 0.3).requires_grad_(True)
        drop = torch.from_numpy(g[name + "_drop"]).view(2, c, 1, 1) if name + "_drop" in g.files else None
        y = O.non_bottleneck_1d(x, sd, "blk", d, dropprob=p if (use_dropout and drop is not None) else 0.0, training=True, drop_scale=drop,
                                use_batchnorm=False)
        y.backward(golden_input(tuple(y.shape), 87) - 0.5)
        np.testing.assert_allclose(y.detach().numpy(), g[name + "_y"], **TOL)
        np.testing.assert_allclose(x.grad.numpy(), g[name + "_dx"], rtol=1e-3, atol=1e-5)
        np.testing.assert_allclose(sd["blk.conv1x3_1.bias"].grad.numpy(), g[name + "_db_conv1x3_1"], rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(sd["blk.conv3x1_1.weight"].grad.numpy(), g[name + "_dw_conv3x1_1"], rtol=1e-3, atol=1e-4)
    full = O.erfnet_param_shapes(20)
    for table, fn, seed, gseed, pref in ((NOBN_DOWN, O.downsampler_block, 14, 88, {3: "encoder.initial_block.", 16: "encoder.layers.0.", 64: "encoder.layers.6."}),
                                         (NOBN_UP, O.upsampler_block, 15, 89, {128: "decoder.layers.0.", 64: "decoder.layers.3."})):
        for name, (ci, co, hw) in table.items():
            shapes = {k[len(pref[ci]):]: s for k, s in full.items() if k.startswith(pref[ci])}
            sd = leaves(shapes, seed)
            x = (golden_input((2, ci) + hw, 50 + ci) - 0.3).requires_grad_(True)
            y = fn(x, sd, "blk", training=True, use_batchnorm=False)
            y.backward(golden_input(tuple(y.shape), gseed) - 0.5)
            np.testing.assert_allclose(y.detach().numpy(), g[name + "_y"], **TOL)
            np.testing.assert_allclose(x.grad.numpy(), g[name + "_dx"], rtol=1e-3, atol=1e-5)
            np.testing.assert_allclose(sd["blk.conv.weight"].grad.numpy(), g[name + "_dw"], rtol=1e-3, atol=1e-4)
            np.testing.assert_allclose(sd["blk.conv.bias"].grad.numpy(), g[name + "_db"], rtol=1e-3, atol=1e-4)


def test_ladder_net(golden_dir):
    """Oracle restatement of train/net_regression_ladder.Net (softmax_classes = 0) vs the reference's outputs."""
    g = np.load(os.path.join(golden_dir, "ladder_net.npz"))
    lines = open(os.path.join(golden_dir, "ladder_net_keys.txt")).read().strip().split("\n")
    assert len(lines) == 395
    shapes = {l.split(" ", 1)[0]: eval(l.split(" ", 1)[1]) for l in lines}
    x = golden_input((2, 3, 64, 128), 71)
    sd = O.make_state_dict(shapes, seed=31)
    with torch.no_grad():
        np.testing.assert_allclose(O.regression_ladder_forward(x, sd).numpy(), g["reg_out"], rtol=1e-4, atol=1e-4)
    shapes["decoder.scalar_output_conv.weight"], shapes["decoder.scalar_output_conv.bias"] = (16, 2, 2, 2), (2,)
    sd2 = O.make_state_dict(shapes, seed=31)
    with torch.no_grad():
        np.testing.assert_allclose(O.regression_ladder_forward(x, sd2).numpy(), g["lik_out"], rtol=1e-4, atol=1e-4)
