"""GPU idle time inside one graph-replayed training step: kernel intervals from CUPTI (torch.profiler chrome trace), merged; reports the busy time,
the idle gaps between kernels and the kernels that run right before the largest gaps.  Guidance only (a profiler is attached)."""
import json
import os
import sys
import tempfile

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NC, synth_batch  # noqa: E402
from erfnet_pytorch_b200 import erfnet as E  # noqa: E402
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
for _ in range(5):
    tr.train_step(x, y)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        tr.train_step(x, y)
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "t.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
# keep the middle step: split by the largest two gaps (between replays)
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print(f"{len(ev)} GPU activities over {(t1 - t0) / 1e3:.3f} ms (3 steps)")
busy, gaps, cur_end, prev = 0.0, [], None, None
for e in ev:
    s, d = e["ts"], e["dur"]
    if cur_end is None:
        cur_end = s
    if s > cur_end:
        gaps.append((s - cur_end, prev["name"][:60] if prev else "", e["name"][:60]))
        busy += 0
        cur_end = s
    if s + d > cur_end:
        busy += s + d - max(cur_end, s)
        cur_end = s + d
    prev = e
print(f"busy {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms in {len(gaps)} gaps")
gaps.sort(reverse=True)
for g, a, b in gaps[:25]:
    print(f"{g:8.1f} us  after {a:60s} before {b}")
import collections
by = collections.Counter()
for g, a, b in gaps:
    by[a.split('(')[0][:50]] += g
print("idle time by preceding kernel:")
for k, v in by.most_common(12):
    print(f"{v:8.1f} us  {k}")
