"""Shared body of the ladder-net tests (train/net_regression_ladder.py, BASELINE config 4): our Net on ``device`` against
tests/golden/ladder_net.npz (outputs, losses and gradients of the reference's own Net).  CPU: kernels emulated; GPU: real kernels."""
import os

import numpy as np
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from erfnet_pytorch_b200 import net_regression_ladder as NL
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input
from tests.test_host_logic import assert_rel

GRAD_KEYS = ("encoder.initial_block.conv.weight", "encoder.block2.5.conv3x1_2.weight", "decoder.ladder_block_1.conv.weight",
             "decoder.ladder_block_2.layer.conv1x3_1.bias", "decoder.scalar_output_conv.weight", "decoder.scalar_decoder_2.layers.0.conv.bias")


def ladder_inputs():
    x = golden_input((2, 3, 64, 128), 71)
    tgt = golden_input((2, 64, 128), 72) * 2.0
    tgt[golden_input((2, 64, 128), 73) < 0.7] = -1.0
    return x, tgt


def make_net(name, device, precision):
    m = NL.Net(likelihood_loss=(name == "lik"))
    m.precision = precision
    sd = O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=31)
    m.load_state_dict(sd, strict=True)
    return m.to(device), sd


def check_ladder_net(golden_dir, device, precision="fp32", out_tol=1e-4, grad_tol=2e-3):
    g = np.load(os.path.join(golden_dir, "ladder_net.npz"))
    keys = [l.split(" ", 1)[0] for l in open(os.path.join(golden_dir, "ladder_net_keys.txt")).read().strip().split("\n")]
    x, tgt = ladder_inputs()
    for name in ("reg", "lik"):
        m, _ = make_net(name, device, precision)
        if name == "reg":
            assert list(m.state_dict().keys()) == keys          # checkpoint layout of the reference net
        m.train()
        y = m(x.to(device))
        assert_rel(y.detach().cpu().numpy(), g[name + "_out"], out_tol, name + " output")
        if name == "reg":
            loss = EV.L1LossMasked()(y, tgt.to(device))
        else:
            loss = EV.LogLikelihoodLossMasked(opt_eps=0.05)(y[:, 0], y[:, 1].pow(2), tgt.to(device))   # as in the golden generator
        loss.backward()
        assert abs(loss.item() - g[name + "_loss"]) <= max(out_tol * 10, 1e-4) * abs(g[name + "_loss"]), (loss.item(), g[name + "_loss"])
        for k in GRAD_KEYS:
            mod, attr = k.rsplit(".", 1)
            got = getattr(m.get_submodule(mod), attr).grad
            assert got is not None, k
            # the likelihood loss is heavy-tailed (a few low-variance pixels carry most of the gradient), so one ReLU mask that flips on
            # a 1e-7 difference moves a gradient by ~1 %: 10x the tolerance of the L1 case
            assert_rel(got.cpu().numpy(), g[f"{name}_g_{k}"], grad_tol * (10 if name == "lik" else 1), f"{name} grad {k}")
        m.eval()
        with torch.no_grad():
            assert_rel(m(x.to(device)).cpu().numpy(), g[name + "_out"], out_tol, name + " eval output")


def check_ladder_net_classes(golden_dir, device, precision="fp32", out_tol=1e-4, grad_tol=2e-3):
    """softmax_classes = 6: SoftMaxConv head, class_power parameters, class-probability losses (dropout forced off as in the golden run)."""
    g = np.load(os.path.join(golden_dir, "ladder_net.npz"))
    x, tgt = ladder_inputs()
    for name in ("cls", "clslik"):
        m = NL.Net(softmax_classes=6, likelihood_loss=(name == "clslik"), spread_class_power=True, late_dropout_prob=0.0)
        m.precision = precision
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout2d):
                mod.p = 0.0
        sd = O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=33)
        sd["class_power"] = torch.linspace(0.0, 2.0, 6).view(1, 6, 1, 1)
        if "class_power_var" in sd:
            sd["class_power_var"] = torch.linspace(0.3, 0.8, 6).view(1, 6, 1, 1)
        m.load_state_dict(sd, strict=True)
        m.to(device).train()
        out = m(x.to(device))
        if name == "cls":
            prob, power = out
            loss = EV.L1LossClassProbMasked()(prob, power, tgt.to(device))
        else:
            prob, power, var = out
            loss = EV.LogLikelihoodLossClassProbMasked(opt_eps=0.05)(prob, power, var, tgt.to(device))
        assert_rel(prob.detach().cpu().numpy(), g[name + "_prob"], out_tol, name + " probabilities")
        np.testing.assert_allclose(prob.detach().sum(1).cpu().numpy(), 1.0, rtol=1e-5)
        loss.backward()
        assert abs(loss.item() - g[name + "_loss"]) <= max(out_tol * 10, 1e-4) * abs(g[name + "_loss"]), (loss.item(), g[name + "_loss"])
        keys = ["class_power", "decoder.scalar_output_conv.convolution.weight", "decoder.ladder_block_1.conv.weight", "encoder.block1.3.conv1x3_1.weight"]
        if name == "clslik":
            keys.append("class_power_var")
        for k in keys:
            if "." in k:
                mod, attr = k.rsplit(".", 1)
                got = getattr(m.get_submodule(mod), attr).grad
            else:
                got = getattr(m, k).grad
            assert got is not None, k
            assert_rel(got.cpu().numpy(), g[f"{name}_g_{k}"], grad_tol * 5, f"{name} grad {k}")
        m.eval()
        with torch.no_grad():
            assert_rel(m(x.to(device))[0].cpu().numpy(), g[name + "_prob"], out_tol, name + " eval probabilities")
