"""In-situ timeline of every conv1d_wgrad_tc launch of ONE graph-replayed training step (batch 6, 512x1024), like tools/trace_step_conv1d.py:
per launch the start relative to the previous weight-gradient launch's end, the wall duration (first CTA entry -> last CTA exit) and the mean
phase lengths of its CTAs."""
import ctypes as C
import os
import sys

import torch

os.environ["ERF_TIMELINE"] = "1"   # the instrumented build of the library (erfnet_pytorch_b200/_lib.py)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import NC, synth_batch  # noqa: E402
from erfnet_pytorch_b200 import _lib  # noqa: E402
from erfnet_pytorch_b200 import erfnet as E  # noqa: E402
from erfnet_pytorch_b200 import synthetic as O  # noqa: E402
from erfnet_pytorch_b200.parallel import FlatAdamTrainer  # noqa: E402

B, LAUNCHES, ROWS, COLS = 6, 68, 300, 10
L = _lib.lib()
L.erf_conv1d_wgrad_tc_set_debug.argtypes = [C.c_void_p, C.c_int]
L.erf_conv1d_wgrad_tc_set_debug.restype = C.c_int
model = E.ERFNet(NC)
model.load_state_dict(O.synthetic_state_dict(model, 0))
model = model.cuda()
tr = FlatAdamTrainer(model, weight=torch.linspace(0.5, 1.5, NC))
pool = [tuple(t.cuda() for t in synth_batch(torch, i, B)) for i in range(4)]
tr.train_step(*pool[0])
buf = torch.zeros(LAUNCHES * ROWS * COLS, device="cuda", dtype=torch.int64)
L.erf_conv1d_wgrad_tc_set_debug(C.c_void_p(buf.data_ptr()), LAUNCHES)
tr.capture_cuda_graph((B, 3, 512, 1024), (B, 512, 1024))
for i in range(6):
    tr.train_step(*pool[i % 4])
torch.cuda.synchronize()
buf.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
tr.train_step(*pool[2])
e1.record()
torch.cuda.synchronize()
L.erf_conv1d_wgrad_tc_set_debug(None, 1)
t = buf.view(LAUNCHES, ROWS, COLS).cpu().double()
ghz = 1.965
print(f"step {e0.elapsed_time(e1):.3f} ms (instrumented replay)")
rows, prev_end, t_first = [], None, None
for k in range(LAUNCHES):
    meta = t[k, ROWS - 1]
    c, hor, dil, bulk, _, units, grid = (int(meta[i]) for i in range(7))
    if c == 0:
        continue
    s = t[k, :grid]
    s = s[s[:, 4] > 0]            # CTAs that had units
    start, end = s[:, 8].min(), s[:, 9].max()
    t_first = start if t_first is None else t_first
    d = lambda a, b: float(((s[:, b] - s[:, a]) / ghz / 1e3).mean())
    rows.append(dict(k=k, c=c, hor=hor, dil=dil, bulk=bulk, upc=units / grid, start=(start - t_first) / 1e3, dur=(end - start) / 1e3,
                     gap=(start - prev_end) / 1e3 if prev_end is not None else 0.0, setup=d(0, 1), first=d(1, 2), loop=d(2, 3), drain=d(3, 4),
                     stage=d(4, 5), red=d(5, 6), exit=d(6, 7), cta=d(0, 7), spread=float((s[:, 8].max() - s[:, 8].min()) / 1e3)))
    prev_end = end
print("  k   C dir dil bulk units/CTA start(us) dur(us) since-prev-end | setup->wait  ->firstMMA   loop  ->acc done  transpose+issue  reduce wait  ->exit | CTA mean | entry spread")
for r in rows:
    print(f"{r['k']:3d} {r['c']:3d}  {'H' if r['hor'] else 'V'} {r['dil']:3d}   {r['bulk']}    {r['upc']:6.1f} {r['start']:9.1f} {r['dur']:7.1f} {r['gap']:8.1f}       | "
          f"{r['setup']:8.2f} {r['first']:10.2f} {r['loop']:7.2f} {r['drain']:9.2f} {r['stage']:12.2f} {r['red']:12.2f} {r['exit']:8.2f} | {r['cta']:7.2f} | {r['spread']:6.2f}")
for c in (64, 128, 16):
    sel = [r for r in rows if r["c"] == c]
    if sel:
        n = len(sel)
        avg = lambda key: sum(r[key] for r in sel) / n
        print(f"C={c:3d}: {n} launches, wall {avg('dur'):6.2f} us avg (sum {sum(r['dur'] for r in sel) / 1e3:.3f} ms) | setup {avg('setup'):5.2f} first {avg('first'):5.2f} loop {avg('loop'):5.2f} "
              f"acc-done {avg('drain'):5.2f} transpose+issue {avg('stage'):5.2f} reduce-wait {avg('red'):5.2f} exit {avg('exit'):5.2f} CTA {avg('cta'):5.2f}")
