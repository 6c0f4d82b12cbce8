import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from erfnet_pytorch_b200 import erfnet as E, evaluation_functions as EV, net_regression_small as NS
from erfnet_pytorch_b200.parallel import FlatAdamTrainer
from oracle import erfnet_oracle as O
crit = EV.L1LossMasked()
name = sys.argv[1]
try:
    x = torch.rand(2, 3, 128, 256, device="cuda"); tgt = torch.rand(2, 128, 256, device="cuda") * 2
    lab = torch.randint(0, 19, (2, 128, 256), device="cuda")
    if name.startswith("fork"):
        net = NS.Net()
        net.load_state_dict(O.make_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed=31))
        net = net.cuda().train()
        if name == "fork_prior_nograd":
            with torch.no_grad():
                net(x)
        if name == "fork_prior_eval":
            net.eval()
            with torch.no_grad():
                net(x)
            net.train()
        tr = FlatAdamTrainer(net, loss_fn=lambda out, t: crit(out, t))
        if name == "fork_trainer_then_eager":
            tr.train_step(x, tgt)
        tr.capture_cuda_graph(tuple(x.shape), tuple(tgt.shape), labels_dtype=torch.float32)
        print(name, "OK", float(tr.train_step(x, tgt)))
    else:
        m = E.ERFNet(20)
        m.load_state_dict(O.make_state_dict(O.erfnet_param_shapes(20), 0))
        m = m.cuda().train()
        if name == "ce_prior_fwd":
            y = m(x)
        tr = FlatAdamTrainer(m)
        tr.capture_cuda_graph(tuple(x.shape), tuple(lab.shape))
        print(name, "OK", float(tr.train_step(x, lab)))
except Exception as e:
    print(name, "FAILED", type(e).__name__, str(e)[:120])
    os._exit(0)
