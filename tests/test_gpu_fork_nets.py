"""GPU: the fork's remaining five net_*.py graphs through the C ABI against the reference's golden outputs / losses / gradients (fp32
check mode), and one of them in bf16 at a tensor-core-eligible size against the fp32 run of the same kernels' check mode."""
import pytest
import torch

from oracle.make_golden import golden_input
from tests.forknet_cases import build_fork_net, check_fork_nets
from tests.test_host_logic import assert_rel

pytestmark = pytest.mark.gpu


def test_fork_nets_fp32_golden(golden_dir):
    check_fork_nets(golden_dir, "cuda", "fp32", out_tol=1e-4, grad_tol=2e-3)


@pytest.mark.parametrize("name", ["regression_small", "classification_ladder"])
def test_fork_net_bf16_tensor_cores(name):
    """1 x 3 x 128 x 1024: every 64/128-channel block, the ladder convs and the head run on tcgen05; compared with the fp32 check mode."""
    x = golden_input((1, 3, 128, 1024), 95).cuda()
    outs = {}
    for precision in ("fp32", "bf16"):
        m, _ = build_fork_net(name, "cuda", precision)
        m.eval()
        with torch.no_grad():
            o = m(x)
        outs[precision] = (o[0] if isinstance(o, tuple) else o).float().cpu().numpy()
    assert_rel(outs["bf16"], outs["fp32"], 3e-2, name + " bf16 vs fp32")


@pytest.mark.parametrize("graph", [False, True])
def test_flat_trainer_generic_loss(graph):
    """The flat trainer (gradient sink, one Adam kernel, optional CUDA graph) on a fork net + masked L1 == autograd + torch.optim.Adam."""
    from tests.flat_generic_cases import check_flat_generic
    check_flat_generic("cuda", "fp32", graph=graph, tol=2e-4)
    check_flat_generic("cuda", "bf16", shape=(1, 3, 128, 1024), graph=graph, tol=5e-2)


def test_generic_trainer_host_fed_stream_keeps_float_targets():
    """train_stream() stages host batches on a side stream; the staging buffers must keep the targets' dtype (float regression targets)."""
    import numpy as np
    from erfnet_pytorch_b200 import evaluation_functions as EV
    from erfnet_pytorch_b200 import net_regression_small
    from erfnet_pytorch_b200.parallel import FlatAdamTrainer
    from oracle import erfnet_oracle as O
    crit = EV.L1LossMasked()
    batches = []
    for i in range(3):
        x = golden_input((1, 3, 64, 128), 150 + i)
        t = golden_input((1, 64, 128), 160 + i) * 2.0
        t[golden_input((1, 64, 128), 170 + i) < 0.5] = -1.0
        batches.append((x.pin_memory(), t.pin_memory()))
    losses = []
    for streamed in (True, False):
        m = net_regression_small.Net()
        m.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, seed=41))
        m.precision = "fp32"
        tr = FlatAdamTrainer(m.cuda(), loss_fn=lambda out, tgt: crit(out, tgt))
        if streamed:
            losses.append([float(l) for l in tr.train_stream(batches)])
        else:
            losses.append([float(tr.train_step(x.cuda(), t.cuda())) for x, t in batches])
    np.testing.assert_allclose(losses[0], losses[1], rtol=1e-5)
    assert losses[0][0] > 0.1          # a truncated (int64) target would change the loss drastically
