#!/usr/bin/env python3
"""Per-kernel totals of an ncu launch list (--metrics gpu__time_duration.sum --csv): index-load kernels apart, the rest
with their share of the device time.  usage: launch_summary.py <launches.csv> [title line]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ki, mi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
    try:
        v = float(r[mi].replace(",", ""))
    except ValueError:
        continue
    ms = v / 1e6 if r[ui] in ("ns", "nsecond") else (v / 1e3 if r[ui] in ("us", "usecond") else v)
    a = agg.setdefault(name, [0, 0.0, 0.0]); a[0] += 1; a[1] += ms; a[2] = max(a[2], ms)
load = [k for k in agg if k.startswith("k_build") or k.startswith("k_resolve_c<0, 1>") or "k_mark" in k or k.startswith("k_synth")]
print("# %s" % (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]))
print("# ncu --metrics gpu__time_duration.sum --clock-control none: cold-cache, serialised -- compare shares, not absolutes\n")
print("index-load / generator kernels (once per process, not in any timed region):")
for k in load:
    print("  %-60s n=%4d total %10.3f ms" % (k, agg[k][0], agg[k][1]))
rest = {k: v for k, v in agg.items() if k not in load}
t2 = sum(v[1] for v in rest.values())
print("\nclassification / text kernels: total %.3f ms" % t2)
for k, v in sorted(rest.items(), key=lambda kv: -kv[1][1]):
    print("  %-60s n=%4d total %10.3f ms share %5.1f%% avg %8.3f ms max %8.3f ms" % (k, v[0], v[1], 100 * v[1] / t2, v[1] / v[0], v[2]))
