"""Turn the ncu CSV of `tools/bench_elementwise.py --once` (metrics: gpu__time_duration.sum, dram__bytes_read/write.sum,
dram__throughput pct) into a markdown table of our (b200::) kernels.  Usage: python tools/hbm_csv_summary.py <csv> [peak_GBs]"""
import csv
import sys
from collections import OrderedDict

path = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 6485.5
rows = OrderedDict()
with open(path) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    k = r["Kernel Name"]
    if "b200::" not in k and not k.startswith("b200"):
        continue
    e = rows.setdefault(r["ID"], {"kernel": k.replace("void ", "").split("(")[0], "grid": r["Grid Size"], "block": r["Block Size"]})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    m = r["Metric Name"]
    if m == "gpu__time_duration.sum":
        e["us"] = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(u, 1.0)
    elif m.startswith("dram__bytes"):
        e[m] = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    elif m.startswith("dram__throughput"):
        e["pct"] = v
print("| id | kernel | grid x block | us (ncu, cold) | DRAM read MB | DRAM write MB | DRAM GB/s | % of measured copy peak "
      f"({peak:.0f} GB/s) | ncu dram__throughput % |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|")
for i, e in rows.items():
    rd, wr = e.get("dram__bytes_read.sum", 0.0), e.get("dram__bytes_write.sum", 0.0)
    gbs = (rd + wr) / e["us"] * 1e3 if e.get("us") else 0.0
    print(f"| {i} | `{e['kernel']}` | {e['grid']} x {e['block']} | {e['us']:.1f} | {rd:.1f} | {wr:.1f} | {gbs:.0f} | "
          f"{100 * gbs / peak:.1f} | {e.get('pct', float('nan')):.1f} |")
