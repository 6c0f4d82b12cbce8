"""Small fixed workload for ncu: 1 warm-up + 1 measured training step and 1 + 1 eval forwards at the bench shape."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NC, synth_batch  # noqa: E402
from erfnet_pytorch_b200 import erfnet as E  # noqa: E402
from erfnet_pytorch_b200.parallel import FlatAdamTrainer  # noqa: E402
from erfnet_pytorch_b200 import synthetic as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
model = E.ERFNet(NC)
model.load_state_dict(O.synthetic_state_dict(model, 0))
model = model.cuda()
x, y = synth_batch(torch, 0, B)
x, y = x.cuda(), y.cuda()
tr = FlatAdamTrainer(model, weight=torch.linspace(0.5, 1.5, NC))
for _ in range(2):
    tr.train_step(x, y)
model.eval()
for _ in range(2):
    with torch.no_grad():
        model(x)
torch.cuda.synchronize()
print("ok")
