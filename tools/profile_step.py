"""Per-kernel CUDA-event breakdown of one forward (eval) and one training step at the bench workload."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NC, synth_batch  # noqa: E402
from erfnet_pytorch_b200 import erfnet as E  # noqa: E402
from erfnet_pytorch_b200 import ops  # noqa: E402
from erfnet_pytorch_b200.parallel import FlatAdamTrainer  # noqa: E402
from erfnet_pytorch_b200 import synthetic as O  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 6
ops.DETAIL = True
model = E.ERFNet(NC)
model.load_state_dict(O.synthetic_state_dict(model, 0))
model = model.cuda()
x, y = synth_batch(torch, 0, B)
x, y = x.cuda(), y.cuda()
tr = FlatAdamTrainer(model, weight=torch.linspace(0.5, 1.5, NC))


def table(title, kt):
    tot = sum(v["ms"] for v in kt.values())
    print(f"--- {title}: {tot:.3f} ms in timed kernels")
    for k, v in sorted(kt.items(), key=lambda kv: -kv[1]["ms"]):
        tf = f"{v['flops'] / (v['ms'] * 1e-3) / 1e12:8.1f} TF/s" if v["flops"] else ""
        print(f"{k:52s} n={v['n']:4d} {v['ms']:9.3f} ms {100 * v['ms'] / tot:5.1f}% {tf}")


model.eval()
for _ in range(2):
    with torch.no_grad():
        model(x)
torch.cuda.synchronize()
torch.cuda._sleep(int(0.02 * 1.9e9))
ops.start_kernel_timing()
with torch.no_grad():
    model(x)
table(f"forward eval bf16 batch {B}", ops.stop_kernel_timing())
for _ in range(2):
    tr.train_step(x, y)
torch.cuda.synchronize()
torch.cuda._sleep(int(0.06 * 1.9e9))     # park the GPU so that the whole step is queued: the event intervals then measure kernels, not the CPU launch path
ops.start_kernel_timing()
tr.train_step(x, y)
table(f"train step bf16 batch {B}", ops.stop_kernel_timing())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
model.eval()
torch.cuda.synchronize()
e0.record()
for _ in range(5):
    with torch.no_grad():
        model(x)
e1.record()
torch.cuda.synchronize()
print(f"forward wall (events) {e0.elapsed_time(e1) / 5:.3f} ms/batch")
