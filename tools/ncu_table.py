"""Key metrics per launch of an .ncu-rep (read with `ncu -i ... --page raw --csv`) as a markdown table."""
import csv
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "dram read MB"), ("dram__bytes_write.sum", "dram write MB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("launch__registers_per_thread", "regs"), ("launch__shared_mem_per_block_dynamic", "dyn smem KB"), ("launch__grid_size", "grid")]
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
ki = h.index("Kernel Name")
cols = [(h.index(m), t) for m, t in WANT if m in h]
print("| kernel | " + " | ".join(t for _, t in cols) + " |")
print("|---|" + "---:|" * len(cols))
for r in rows[2:]:
    name = r[ki].split("(")[0].replace("void ", "").replace("erfk::", "")
    vals = []
    for i, t in cols:
        v = float(r[i].replace(",", "")) if r[i] else 0.0
        u = units[i]
        if "MB" in t:
            v = v / 1e6 if u == "byte" else (v / 1e3 if u == "Kbyte" else (v if u == "Mbyte" else v * 1e3))
        if t == "time us":
            v = v / 1e3 if u in ("ns", "nsecond") else v
        vals.append(f"{v:.2f}" if not float(v).is_integer() else f"{int(v)}")
    print(f"| `{name[:48]}` | " + " | ".join(vals) + " |")
