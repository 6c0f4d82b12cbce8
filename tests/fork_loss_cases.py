"""Shared body of the fork-loss parity tests: runs erfnet_pytorch_b200.evaluation_functions on ``device`` against
tests/golden/fork_losses.npz (outputs of the reference's own classes).  Used with the emulated kernels on the CPU
(host logic) and with the real kernels on the GPU."""
import os

import numpy as np
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from oracle.make_golden import fork_loss_inputs


def check_fork_losses(golden_dir, device, rtol=1e-4):
    z = np.load(os.path.join(golden_dir, "fork_losses.npz"))
    d = {k: v.to(device) for k, v in fork_loss_inputs().items()}
    cases = {
        "l1": (EV.L1LossMasked(), (d["out"], d["target"]), (0,)),
        "loglik": (EV.LogLikelihoodLossMasked(opt_eps=1e-4), (d["out"][:, 0], d["var"], d["target"]), (0, 1)),
        "l1cp": (EV.L1LossClassProbMasked(), (d["prob"], d["cost"], d["target"]), (0, 1)),
        "loglikcp": (EV.LogLikelihoodLossClassProbMasked(opt_eps=1e-4), (d["prob"], d["cost"], d["cvar"], d["target"]), (0, 1, 2)),
        "mse": (lambda a, b: EV.MSELossWeighted(3.0)(a, b, 0.25), (d["out"], d["teacher"]), (0, 1)),
    }
    for name, (crit, args, gidx) in cases.items():
        leaves = [a.clone().requires_grad_(i in gidx) for i, a in enumerate(args)]
        loss = crit(*leaves)
        (loss * 1.0).backward()
        np.testing.assert_allclose(loss.item(), z[name + "_loss"], rtol=1e-5, err_msg=name)
        for i in gidx:
            g = leaves[i].grad
            assert g is not None and g.shape == leaves[i].shape, f"{name} grad {i}"
            np.testing.assert_allclose(g.cpu().numpy(), z[f"{name}_g{i}"], rtol=rtol, atol=1e-7, err_msg=f"{name} grad {i}")
    # upstream gradient scaling (go != 1)
    o = d["out"].clone().requires_grad_(True)
    (EV.L1LossMasked()(o, d["target"]) * 2.5).backward()
    np.testing.assert_allclose(o.grad.cpu().numpy(), 2.5 * z["l1_g0"], rtol=rtol, atol=1e-7)
    acc = EV.ClassificationAccuracy()(d["prob"], d["labels"])
    macc, cacc = EV.MeanAccuracy(d["prob"].shape[1])(d["prob"], d["labels"])
    np.testing.assert_allclose(float(acc), z["acc"], rtol=1e-6)
    np.testing.assert_allclose(float(macc), z["mean_acc"], rtol=1e-6)
    np.testing.assert_allclose(cacc, z["class_acc"], rtol=1e-6)
