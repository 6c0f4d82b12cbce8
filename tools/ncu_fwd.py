"""Eval-mode forwards only (for ncu captures of the inference kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NC, synth_batch  # noqa: E402
from erfnet_pytorch_b200 import erfnet as E  # noqa: E402
from erfnet_pytorch_b200 import synthetic as O  # noqa: E402

model = E.ERFNet(NC)
model.load_state_dict(O.synthetic_state_dict(model, 0))
model = model.cuda().eval()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
x, _ = synth_batch(torch, 0, B)
x = x.cuda()
for _ in range(3):
    with torch.no_grad():
        model(x)
torch.cuda.synchronize()
print("ok")
