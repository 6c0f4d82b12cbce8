"""GPU: masked regression losses and accuracy metrics (train/evaluation_functions.py:32-163) through the C ABI against the
reference's golden outputs, plus size-independent properties at the bench size (6 x 512 x 1024)."""
import numpy as np
import pytest
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from oracle import erfnet_oracle as O
from tests.fork_loss_cases import check_fork_losses

pytestmark = pytest.mark.gpu


def test_fork_losses_golden(golden_dir):
    check_fork_losses(golden_dir, "cuda")


def test_fork_losses_full_size_properties():
    g = torch.Generator().manual_seed(5)
    b, c, h, w = 6, 6, 512, 1024
    out = (torch.rand(b, 1, h, w, generator=g) * 2).cuda()
    tgt = torch.rand(b, h, w, generator=g) * 2
    tgt[torch.rand(b, h, w, generator=g) < 0.7] = -1.0
    tgt = tgt.cuda()
    n_valid = int((tgt != -1).sum())
    # oracle formula evaluated by torch on the GPU in fp64 (same arithmetic, bigger accumulator)
    ref = O.l1_loss_masked(out.double(), tgt.double()).item()
    o = out.clone().requires_grad_(True)
    loss = EV.L1LossMasked()(o, tgt)
    loss.backward()
    assert abs(loss.item() - ref) <= 1e-6 * abs(ref)
    gr = o.grad[:, 0]
    assert torch.count_nonzero(gr[tgt == -1]) == 0                       # unlabelled pixels get no gradient
    assert torch.all(gr[tgt != -1].abs() * n_valid == 1.0) or torch.allclose(gr[tgt != -1].abs() * n_valid, torch.ones((), device="cuda"))
    # linearity in the scale factor and symmetry of the consistency loss
    a, t2 = out, (torch.rand(b, 1, h, w, generator=g) * 2).cuda()
    l1 = EV.MSELossWeighted(1.0)(a, t2, 1.0).item()
    l2 = EV.MSELossWeighted(4.0)(t2, a, 0.5).item()
    assert abs(l2 - 2 * l1) <= 1e-6 * abs(l1)
    # class-probability loss with a one-hot probability map equals the plain L1 loss against that class's cost
    cost = torch.linspace(0.1, 1.9, c).view(1, c, 1, 1).cuda()
    cls = torch.randint(0, c, (b, h, w), generator=g).cuda()
    prob = torch.nn.functional.one_hot(cls, c).permute(0, 3, 1, 2).float().contiguous()
    lcp = EV.L1LossClassProbMasked()(prob, cost, tgt).item()
    lpl = EV.L1LossMasked()(cost.view(-1)[cls].unsqueeze(1), tgt).item()
    assert abs(lcp - lpl) <= 1e-6 * abs(lpl)
    # accuracy: predictions == labels on labelled pixels -> 1.0; counts are integers from the confusion histogram
    labels = cls.clone()
    labels[tgt == -1] = -1
    assert float(EV.ClassificationAccuracy()(prob, labels)) == 1.0
    macc, cacc = EV.MeanAccuracy(c)(prob, labels)
    assert float(macc) == 1.0 and np.all(cacc == 1.0)
    shifted = (cls + 1) % c
    assert float(EV.ClassificationAccuracy()(prob, torch.where(labels >= 0, shifted, labels))) == 0.0


def test_fork_loss_errors():
    with pytest.raises(RuntimeError):
        EV.L1LossMasked()(torch.zeros(1, 1, 4, 4), torch.zeros(1, 4, 4))  # CPU tensors: no fallback
    with pytest.raises(RuntimeError):
        EV.L1LossClassProbMasked()(torch.zeros(1, 3, 4, 4).cuda(), torch.zeros(1, 4, 1, 1).cuda(), torch.zeros(1, 4, 4).cuda())
