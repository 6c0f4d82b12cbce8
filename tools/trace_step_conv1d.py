"""In-situ timeline of every conv1d_tc launch of ONE graph-replayed training step at the bench workload (batch 6, 512x1024): the
instrumented kernel (erf_conv1d_tc_set_debug) records per-CTA SM-clock stamps of its phases plus globaltimer stamps at entry / exit, so
the table shows, per launch, when it started relative to its predecessor's end (gaps, PDL overlap), its wall duration (first CTA entry
-> last CTA exit) and the mean phase lengths -- i.e. where the time of the dominant kernel goes INSIDE the step, not in isolation."""
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

B, LAUNCHES, ROWS, COLS = 6, 136, 300, 12
L = _lib.lib()
L.erf_conv1d_tc_set_debug.argtypes = [C.c_void_p, C.c_int]
L.erf_conv1d_tc_set_debug.restype = C.c_int
model = E.ERFNet(NC)
model.load_state_dict(O.synthetic_state_dict(model, 0))
model = model.cuda()
tr = FlatAdamTrainer(model, weight=torch.linspace(0.5, 1.5, NC))
pool = [tuple(t.cuda() for t in synth_batch(torch, i, B)) for i in range(4)]
tr.train_step(*pool[0])
buf = torch.zeros(LAUNCHES * ROWS * COLS, device="cuda", dtype=torch.int64)
# launch indices are handed out per call, modulo LAUNCHES: the capture's warm-up steps wrap around and the captured step gets 0..135 in order
L.erf_conv1d_tc_set_debug(C.c_void_p(buf.data_ptr()), LAUNCHES)
tr.capture_cuda_graph((B, 3, 512, 1024), (B, 512, 1024))
for i in range(6):
    tr.train_step(*pool[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
buf.zero_()
e0.record()
tr.train_step(*pool[2])
e1.record()
torch.cuda.synchronize()
L.erf_conv1d_tc_set_debug(None, 1)
t = buf.view(LAUNCHES, ROWS, COLS).cpu().double()
ghz = 1.965
print(f"step {e0.elapsed_time(e1):.3f} ms (instrumented replay)")
t_first = None
prev_end = None
rows = []
for k in range(LAUNCHES):
    meta = t[k, ROWS - 1]
    c, hor, dil, aux, red, units, grid = (int(meta[i]) for i in range(7))
    if c == 0:
        continue
    s = t[k, :grid]
    start, end = s[:, 8].min(), s[:, 9].max()
    if t_first is None:
        t_first = start
    d = lambda a, b: float(((s[:, b] - s[:, a]) / ghz / 1e3).mean())
    epi_end = torch.maximum(s[:, 5], s[:, 6])
    rows.append(dict(k=k, c=c, hor=hor, dil=dil, aux=aux, red=red, upc=units / grid, start=(start - t_first) / 1e3, dur=(end - start) / 1e3,
                     gap=(start - prev_end) / 1e3 if prev_end is not None else 0.0, setup=d(0, 1), wts=d(1, 2), first=d(2, 3), loop=d(3, 4),
                     tail=float(((epi_end - s[:, 4]) / ghz / 1e3).mean()), cta=d(0, 7),
                     st_loop=float(((s[:, 10] - s[:, 4]) / ghz / 1e3).mean()) if red else 0.0, st_atom=d(10, 11) if red else 0.0, spread=float((s[:, 8].max() - s[:, 8].min()) / 1e3)))
    prev_end = end
print("  k   C  dir dil aux red units/CTA  start(us)  dur(us) since-prev-end | setup->wait  +weights  +firstMMA  loop   tail | CTA mean | entry spread | stats: lastMMA->loop end, +atomics")
for r in rows:
    print(f"{r['k']:3d} {r['c']:3d}  {'H' if r['hor'] else 'V'}  {r['dil']:3d}  {r['aux']}   {r['red']}   {r['upc']:6.1f}  {r['start']:9.1f} {r['dur']:8.1f} {r['gap']:8.1f}       | "
          f"{r['setup']:8.2f} {r['wts']:9.2f} {r['first']:9.2f} {r['loop']:6.2f} {r['tail']:6.2f} | {r['cta']:7.2f} | {r['spread']:6.2f} | {r['st_loop']:6.2f} {r['st_atom']:6.2f}")
for c in (64, 128, 16):
    sel = [r for r in rows if r["c"] == c]
    if sel:
        n = len(sel)
        print(f"C={c:3d}: {n} launches, wall {sum(r['dur'] for r in sel) / n:6.2f} us avg (sum {sum(r['dur'] for r in sel) / 1e3:.3f} ms) | setup {sum(r['setup'] for r in sel) / n:5.2f} "
              f"weights {sum(r['wts'] for r in sel) / n:5.2f} first {sum(r['first'] for r in sel) / n:5.2f} loop {sum(r['loop'] for r in sel) / n:5.2f} tail {sum(r['tail'] for r in sel) / n:5.2f} "
              f"CTA {sum(r['cta'] for r in sel) / n:5.2f}")
