"""Shared body of the tests of the five remaining fork nets (train/net_regression.py, net_regression_small.py,
net_regression_ladder_small.py, net_classification.py, net_classification_ladder.py): our Net on ``device`` against
tests/golden/fork_nets.npz (state_dict layout, outputs, losses and gradients of the reference's own classes)."""
import importlib
import os

import numpy as np
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from oracle import erfnet_oracle as O
from oracle.make_golden import FORK_NET_GRADS, FORK_NETS, fork_net_inputs, fork_net_loss
from tests.test_host_logic import assert_rel


def build_fork_net(name, device, precision="fp32"):
    modname, kw, kind = FORK_NETS[name]
    m = importlib.import_module("erfnet_pytorch_b200." + modname).Net(**kw)
    m.precision = precision
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout2d):
            mod.p = 0.0
    sd = O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=41)
    if "class_power" in sd:
        sd["class_power"] = torch.linspace(0.2, 1.8, sd["class_power"].numel()).view_as(sd["class_power"])
    m.load_state_dict(sd, strict=True)
    return m.to(device), kind


def check_fork_nets(golden_dir, device, precision="fp32", out_tol=1e-4, grad_tol=2e-3, names=None):
    g = np.load(os.path.join(golden_dir, "fork_nets.npz"))
    x, tgt, labels = fork_net_inputs()
    for name in names or FORK_NETS:
        m, kind = build_fork_net(name, device, precision)
        assert [f"{k} {tuple(v.shape)}" for k, v in m.state_dict().items()] == list(g[name + "_keys"]), name   # checkpoint layout of the reference net
        m.train()
        out = m(x.to(device))
        first = out[0] if isinstance(out, tuple) else out
        assert_rel(first.detach().cpu().numpy(), g[name + "_out"], out_tol, name + " output")
        if isinstance(out, tuple):
            assert out[1] is m.class_power
        loss = fork_net_loss(EV, kind, out, tgt.to(device), labels.to(device))
        loss.backward()
        assert abs(loss.item() - g[name + "_loss"]) <= max(out_tol * 10, 1e-4) * abs(g[name + "_loss"]), (name, loss.item(), g[name + "_loss"])
        for k in FORK_NET_GRADS + (("class_power",) if kind == "prob" else ()):
            got = m.get_parameter(k).grad
            assert got is not None, (name, k)
            # heavy-tailed likelihood gradients: see tests/ladder_cases.py
            assert_rel(got.cpu().numpy(), g[f"{name}_g_{k}"], grad_tol * (10 if kind == "lik" else 5 if kind == "prob" else 1), f"{name} grad {k}")
        assert m.encoder.output_conv.weight.grad is None          # unused in these graphs, as in the reference
        m.eval()
        with torch.no_grad():
            o2 = m(x.to(device))
            assert_rel((o2[0] if isinstance(o2, tuple) else o2).cpu().numpy(), g[name + "_out"], out_tol, name + " eval output")
