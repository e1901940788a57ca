#!/usr/bin/env python3
"""Per-kernel times of the last text span in an ncu launch list (gpu__time_duration.sum CSV)."""
import csv, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = csv.reader(lines); hdr = next(r)
ki, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
seq = []
for row in r:
    v = float(row[vi].replace(",", "")); u = row[ui]
    v = v / 1e6 if u == "ns" else (v / 1e3 if u == "us" else v)
    seq.append((row[ki].split("(")[0].replace("void ", ""), v, row[gi]))
idx = [i for i, (n, _, _) in enumerate(seq) if n.startswith("k_nl_count")]
start = idx[-1]; tot = 0
for n, v, g in seq[start:start + 29]:
    print("%-28s %8.4f ms grid %s" % (n[:28], v, g)); tot += v
print("sum %.3f ms" % tot)
