#!/usr/bin/env python
"""Summarise an ncu --csv launch list (metrics gpu__time_duration.sum [+ dram__bytes_read.sum, dram__bytes_write.sum])
by kernel name:  python scratch/ncu_launch_summary.py gpurun_out/launches.csv[.gz] [--from-id N] > profiles/...txt"""
import csv
import gzip
import re
import sys
from collections import defaultdict

path = sys.argv[1]
lo = int(sys.argv[sys.argv.index("--from-id") + 1]) if "--from-id" in sys.argv else 0
hi = int(sys.argv[sys.argv.index("--to-id") + 1]) if "--to-id" in sys.argv else 1 << 60
op = gzip.open if path.endswith(".gz") else open
rows = defaultdict(lambda: defaultdict(float))
names = {}
with op(path, "rt", errors="replace") as fh:
    rd = csv.reader(l for l in fh if l.startswith('"'))
    for r in rd:
        if len(r) < 15 or r[0] == "ID":
            continue
        i = int(r[0])
        if i < lo or i >= hi:
            continue
        names[i] = r[4]
        rows[i][r[12]] = float(r[14].replace(",", ""))


def short(n):
    n = re.sub(r"\(.*$", "", n)
    n = n.replace("gwbse::<unnamed>::", "").replace("gwbse::", "")
    return n[:110]


agg = defaultdict(lambda: [0.0, 0, 0.0, 0.0])
for i, m in rows.items():
    a = agg[short(names[i])]
    a[0] += m.get("gpu__time_duration.sum", 0.0) / 1e6
    a[1] += 1
    a[2] += m.get("dram__bytes_read.sum", 0.0)
    a[3] += m.get("dram__bytes_write.sum", 0.0)
tot = sum(a[0] for a in agg.values())
print(f"# {path}: {len(rows)} launches, total kernel time {tot:.1f} ms (ncu: cold cache, serialised - compare shares)")
print(f"{'ms':>11} {'share':>7} {'launches':>9} {'avg_ms':>9} {'dram_GB':>9} {'GB/s':>8}  kernel")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    gb = (a[2] + a[3]) / 1e9
    bw = gb / (a[0] / 1e3) if a[0] > 0 else 0.0
    print(f"{a[0]:11.2f} {100 * a[0] / tot:6.1f}% {a[1]:9d} {a[0] / a[1]:9.3f} {gb:9.2f} {bw:8.0f}  {k}")
