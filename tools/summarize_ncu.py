"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total time, share."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
ki, mi, vi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r is hdr or len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("erfk::", "")
    name = re.sub(r"at::native::.*?(\w+_kernel\w*).*", r"torch:\1", name)
    t = float(r[vi].replace(",", ""))
    t = t / 1e3 if r[ui] in ("ns", "nsecond") else (t if r[ui] in ("us", "usecond") else t * 1e3)
    agg[name][0] += 1
    agg[name][1] += t
tot = sum(v[1] for v in agg.values())
print(f"| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"| `{k[:70]}` | {n} | {t:.1f} | {100 * t / tot:.1f}% | {t / n:.1f} |")
print(f"| **total** | {sum(v[0] for v in agg.values())} | {tot:.1f} | | |")
