"""FlatAdamTrainer with a ``loss_fn`` (any net + any criterion, here the fork's small regression net + L1LossMasked, BASELINE config 4 style)
against the plain autograd + torch.optim.Adam loop of train/main.py:213, 350-381 on the same modules."""
import numpy as np
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from erfnet_pytorch_b200 import net_regression_small
from erfnet_pytorch_b200.parallel import FlatAdamTrainer
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input


def _net(device, precision):
    m = net_regression_small.Net()
    m.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=41))
    m.precision = precision
    return m.to(device).train()


def check_flat_generic(device, precision="fp32", shape=(1, 3, 64, 128), graph=False, tol=1e-5):
    x = golden_input(shape, 111).to(device)
    tgt = (golden_input((shape[0],) + shape[2:], 112) * 2.0)
    tgt[golden_input((shape[0],) + shape[2:], 113) < 0.6] = -1.0
    tgt = tgt.to(device)
    crit = EV.L1LossMasked()
    ref = _net(device, precision)
    opt = torch.optim.Adam(ref.parameters(), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
    ref_losses = []
    for _ in range(3):
        opt.zero_grad()
        loss = crit(ref(x), tgt)
        loss.backward()
        opt.step()
        ref_losses.append(float(loss.detach()))
    m = _net(device, precision)
    tr = FlatAdamTrainer(m, loss_fn=lambda out, t: crit(out, t))
    if graph:
        tr.capture_cuda_graph(tuple(x.shape), tuple(tgt.shape), labels_dtype=torch.float32)
    losses = [float(tr.train_step(x, tgt)) for _ in range(3)]
    np.testing.assert_allclose(losses, ref_losses, rtol=max(tol, 1e-5))
    for (k, a), (_, b) in zip(m.named_parameters(), ref.named_parameters()):
        scale = float(b.detach().abs().max()) + 1e-12
        assert float((a.detach() - b.detach()).abs().max()) <= max(tol, 1e-5) * scale, k
    assert m.encoder.output_conv.weight.grad is not None and float(m.encoder.output_conv.weight.grad.abs().max()) == 0.0   # no gradient: slice stays zero
