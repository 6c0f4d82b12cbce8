"""Kernel-level time table (torch.profiler / CUPTI) of the graph-replayed training step and the eval forward -- guidance only,
never a bench number."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NC, synth_batch  # noqa: E402
from erfnet_pytorch_b200 import erfnet as E  # noqa: E402
from erfnet_pytorch_b200.graphs import GraphedInference  # noqa: E402
from erfnet_pytorch_b200.parallel import FlatAdamTrainer  # noqa: E402
from erfnet_pytorch_b200 import synthetic as O  # noqa: E402

B = 6
model = E.ERFNet(NC)
model.load_state_dict(O.synthetic_state_dict(model, 0))
model = model.cuda()
x, y = synth_batch(torch, 0, B)
x, y = x.cuda(), y.cuda()
tr = FlatAdamTrainer(model, weight=torch.linspace(0.5, 1.5, NC))
tr.capture_cuda_graph(x.shape, y.shape)
for _ in range(3):
    tr.train_step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
print("=== training step (graph replay), 3 steps")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
model.eval()
inf = GraphedInference(model, x)
for _ in range(3):
    inf(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        inf(x)
    torch.cuda.synchronize()
print("=== eval forward (graph replay), 3 batches")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))
