"""CPU: the host-side contracts added in round 2 after the advisor's review (kernels emulated by tests/emu.py):
shape / device validation in front of the loss kernels, learning-rate changes that reach a captured optimiser launch, optimiser
checkpoints in torch.optim's layout, detection of orphaned flat buffers, the opt-in strict label check."""
import copy

import numpy as np
import pytest
import torch

from erfnet_pytorch_b200 import evaluation_functions as EV
from erfnet_pytorch_b200 import net_regression_small, ops
from erfnet_pytorch_b200.optim import FlatSGD, MeanTeacher
from erfnet_pytorch_b200.parallel import FlatAdamTrainer
from oracle import erfnet_oracle as O
from oracle.make_golden import golden_input
from tests.emu import emulated_kernels


def test_cross_entropy_rejects_mismatched_shapes():
    """The reference's nn.CrossEntropyLoss raises on a size mismatch (train/evaluation_functions.py:11-14); the kernels would read and
    WRITE out of bounds instead, so the host checks first."""
    crit = EV.CrossEntropyLoss2d()
    logits = torch.zeros(2, 20, 8, 16, requires_grad=True)
    with emulated_kernels():
        crit(logits, torch.zeros(2, 8, 16, dtype=torch.int64))                      # fine
        for bad in (torch.zeros(2, 16, 32, dtype=torch.int64),                       # full-res labels against 1/2-res logits
                    torch.zeros(3, 8, 16, dtype=torch.int64),                        # batch mismatch
                    torch.zeros(2, 1, 8, 16, dtype=torch.int64)):                    # [B,1,H,W] instead of [B,H,W]
            with pytest.raises(RuntimeError, match="size mismatch"):
                crit(logits, bad)
        with pytest.raises(RuntimeError, match="size mismatch"):
            crit(torch.zeros(2, 8, 16, 24), torch.zeros(2, 16, 16, dtype=torch.int64), nhwc_classes=20)
        with pytest.raises(RuntimeError, match="pitch"):
            crit(torch.zeros(2, 8, 16, 16), torch.zeros(2, 8, 16, dtype=torch.int64), nhwc_classes=20)
        with pytest.raises(RuntimeError, match="weight has"):
            EV.CrossEntropyLoss2d(weight=torch.ones(19))(logits, torch.zeros(2, 8, 16, dtype=torch.int64))


def test_cross_entropy_rejects_host_tensors_without_emulation():
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.check_ce_shapes(torch.zeros(1, 20, 4, 4), torch.zeros(1, 4, 4, dtype=torch.int64), None, 20, True)


def test_same_device_check():
    a = torch.zeros(2, 2)
    with emulated_kernels():
        ops._same_device("x", a, target=torch.zeros(2), weight=None)
        with pytest.raises(RuntimeError, match="is on meta"):
            ops._same_device("x", a, target=torch.zeros(2, device="meta"))


def test_masked_losses_reject_mismatched_shapes():
    t = torch.rand(2, 8, 16)
    with emulated_kernels():
        EV.L1LossMasked()(torch.rand(2, 1, 8, 16), t)
        with pytest.raises(RuntimeError, match="size mismatch"):
            EV.L1LossMasked()(torch.rand(2, 1, 4, 8), t)                             # 1/2-res prediction against full-res target
        with pytest.raises(RuntimeError, match="size mismatch"):
            EV.LogLikelihoodLossMasked(1e-4)(torch.rand(2, 8, 16), torch.rand(2, 4, 16), t)
        prob = torch.softmax(torch.rand(2, 5, 8, 16), 1)
        cost = torch.rand(1, 5, 1, 1)
        EV.L1LossClassProbMasked()(prob, cost, t)
        with pytest.raises(RuntimeError, match="size mismatch"):
            EV.L1LossClassProbMasked()(prob[:, :, :4], cost, t)
        with pytest.raises(RuntimeError, match="size mismatch"):
            EV.L1LossClassProbMasked()(prob[:1], cost, t)
        with pytest.raises(RuntimeError, match="per-class"):
            EV.L1LossClassProbMasked()(prob, torch.rand(1, 4, 1, 1), t)


def test_strict_labels_matches_torch_on_out_of_range_targets():
    logits = torch.zeros(1, 20, 4, 4)
    y = torch.zeros(1, 4, 4, dtype=torch.int64)
    y[0, 0, 0] = -1          # ignore_index: fine
    with emulated_kernels():
        EV.CrossEntropyLoss2d(strict_labels=True)(logits, y)
        y[0, 1, 1] = 255     # an un-relabelled Cityscapes void id: torch asserts, the kernels would silently skip it
        with pytest.raises(RuntimeError, match="out of bounds"):
            EV.CrossEntropyLoss2d(strict_labels=True)(logits, y)
        EV.CrossEntropyLoss2d()(logits, y)   # default: documented divergence (skipped like ignore_index)


def _net(seed=41):
    m = net_regression_small.Net()
    m.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=seed))
    m.precision = "fp32"
    return m.train()


def _data():
    x = golden_input((1, 3, 32, 64), 211)
    t = golden_input((1, 32, 64), 212) * 2.0
    t[golden_input((1, 32, 64), 213) < 0.5] = -1.0
    return x, t


def _same_params(a, b, tol=2e-5):
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        scale = float(q.detach().abs().max()) + 1e-12
        assert float((p.detach() - q.detach()).abs().max()) <= tol * scale, k


def test_trainer_checkpoint_is_torch_adam_layout_and_resumes_both_ways():
    """train/main.py:229, 783 save / resume checkpoint['optimizer'] = torch.optim.Adam.state_dict()."""
    crit = EV.L1LossMasked()
    x, t = _data()
    with emulated_kernels():
        # (a) trainer -> torch: two steps here, hand the state to torch.optim.Adam, both continue for two steps
        m = _net()
        tr = FlatAdamTrainer(m, loss_fn=lambda o, tt: crit(o, tt))
        for _ in range(2):
            tr.train_step(x, t)
        sd = tr.state_dict()
        ref = _net()
        ref.load_state_dict(m.state_dict())
        opt = torch.optim.Adam(ref.parameters(), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
        torch_sd = opt.state_dict()
        assert sorted(sd["param_groups"][0]["params"]) == sorted(torch_sd["param_groups"][0]["params"])
        assert set(sd["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
        opt.load_state_dict(sd)
        for _ in range(2):
            tr.train_step(x, t)
            opt.zero_grad()
            crit(ref(x), t).backward()
            opt.step()
        _same_params(m, ref)
        # (b) torch -> trainer: a fresh trainer resumes from torch's optimiser state
        m2 = _net(seed=7)
        m2.load_state_dict(ref.state_dict())
        tr2 = FlatAdamTrainer(m2, loss_fn=lambda o, tt: crit(o, tt), lr=5e-3)
        tr2.load_state_dict(opt.state_dict())
        assert tr2.step_count == 4 and tr2.lr == 1e-4
        tr2.train_step(x, t)
        opt.zero_grad()
        crit(ref(x), t).backward()
        opt.step()
        _same_params(m2, ref)
        # (c) own round trip
        tr3 = FlatAdamTrainer(_net(), loss_fn=lambda o, tt: crit(o, tt))
        tr3.load_state_dict(copy.deepcopy(tr2.state_dict()))
        assert tr3.step_count == tr2.step_count and torch.equal(tr3.flat_m, tr2.flat_m) and torch.equal(tr3.flat_v, tr2.flat_v)


@pytest.mark.parametrize("sgd", [False, True])
def test_learning_rate_changes_reach_the_graph_mode_optimiser(sgd):
    """After capture_cuda_graph the optimiser launch reads lr / weight decay from device memory: trainer.lr = ... (LambdaLR.step() of
    train/main.py:238-239) must change the very next step.  The capture itself needs a GPU (tests/test_gpu_graphs.py); here the
    graph-mode code path (device-side hyper-parameters and step counter) runs eagerly on the emulated kernels."""
    crit = EV.L1LossMasked()
    x, t = _data()
    with emulated_kernels():
        m, ref = _net(), _net()
        if sgd:
            tr = FlatAdamTrainer(m, loss_fn=lambda o, tt: crit(o, tt), lr=0.01, weight_decay=1e-4, sgd_momentum=0.9)
            opt = torch.optim.SGD(ref.parameters(), 0.01, momentum=0.9, weight_decay=1e-4)
        else:
            tr = FlatAdamTrainer(m, loss_fn=lambda o, tt: crit(o, tt))
            opt = torch.optim.Adam(ref.parameters(), 1e-4, (0.9, 0.999), eps=1e-4, weight_decay=1e-6)
        tr._step_dev = torch.zeros(1, dtype=torch.int32)                       # what capture_cuda_graph sets up
        tr._hyper_dev = torch.tensor([tr.lr, tr.wd], dtype=torch.float32)
        for step in range(4):
            if step == 2:
                tr.set_lr(tr.lr * 0.1)
                for g in opt.param_groups:
                    g["lr"] *= 0.1
            tr.train_step(x, t)
            opt.zero_grad()
            crit(ref(x), t).backward()
            opt.step()
        assert abs(float(tr._hyper_dev[0]) - opt.param_groups[0]["lr"]) <= 1e-9
        assert tr.param_groups[0]["lr"] == pytest.approx(opt.param_groups[0]["lr"])
        _same_params(m, ref)


def test_flat_sgd_checkpoint_is_torch_sgd_layout():
    x, t = _data()
    crit = EV.L1LossMasked()
    with emulated_kernels():
        m, ref = _net(), _net()
        a = FlatSGD(m.parameters(), 0.01, momentum=0.9, weight_decay=1e-4)
        b = torch.optim.SGD(ref.parameters(), 0.01, momentum=0.9, weight_decay=1e-4)
        for opt, net in ((a, m), (b, ref)):
            opt.zero_grad()
            crit(net(x), t).backward()
            opt.step()
        sd = a.state_dict()
        assert set(sd["state"][0]) == {"momentum_buffer"} and sd["param_groups"][0]["params"] == list(range(len(a.params)))
        b2 = torch.optim.SGD(ref.parameters(), 0.5, momentum=0.1)
        b2.load_state_dict({k: v for k, v in sd.items() if k != "steps"})   # torch accepts the layout
        a2 = FlatSGD(_net(seed=3).parameters(), 0.5, momentum=0.9)
        a2.load_state_dict(b.state_dict())                                   # and this class accepts torch's
        assert a2.steps == 1 and a2.param_groups[0]["lr"] == 0.01
        want = torch.cat([(b.state[p]["momentum_buffer"] if p in b.state else torch.zeros_like(p)).reshape(-1) for p in ref.parameters()])
        assert torch.equal(a2.flat_buf, want)
        np.testing.assert_allclose(a2.flat_buf.numpy(), a.flat_buf.numpy(), rtol=1e-3, atol=1e-3)   # the two optimisers saw the same gradients


def test_orphaned_flat_buffers_are_detected():
    """MeanTeacher / FlatSGD capture flat buffers at construction; a trainer that takes ownership of the parameters afterwards (or a
    Module.to()) re-allocates them -- updating the captured buffers would average orphans in silence."""
    crit = EV.L1LossMasked()
    with emulated_kernels():
        student, teacher = _net(), _net(seed=5)
        mt = MeanTeacher(student, teacher)
        mt.update(0.5)                                                        # fine
        FlatAdamTrainer(student, loss_fn=lambda o, tt: crit(o, tt))           # re-creates the student's parameters
        with pytest.raises(RuntimeError, match="no longer aliases"):
            mt.update(0.5)
        # the documented order works: trainer first, mean teacher last
        student, teacher = _net(), _net(seed=5)
        FlatAdamTrainer(student, loss_fn=lambda o, tt: crit(o, tt))
        MeanTeacher(student, teacher).update(0.5)
        net = _net()
        opt = FlatSGD(net.parameters(), 0.1, momentum=0.9)
        for p in net.parameters():
            p.data = p.data.clone()                                           # what Module.to() / .cuda() does
        with pytest.raises(RuntimeError, match="no longer aliases"):
            opt.step()
