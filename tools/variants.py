#!/usr/bin/env python3
"""A/B timing of kernel variants on the bench workload (env CFB_GROUP = lanes per walk)."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from centrifuge_b200 import capi
genera = int(os.environ.get("CFB_BENCH_GENERA", 900)); n = int(os.environ.get("CFB_BENCH_READS", 2000000))
base, d = bench.get_index(genera, 10, 1000000, 12345)
codes = bench.make_reads(genera, 10, 1000000, 12345, n, 100, 1000)
lens = np.full(n, 100, dtype=np.uint32); offs = np.arange(n, dtype=np.uint64) * np.uint64(100)
flags = ((codes == 4).sum(axis=1) <= 15).astype(np.uint8)
batch = capi.make_batch(codes.reshape(-1), offs, lens, None, None, flags)
ix = capi.Index(base, 0)
ref = None
for g in sys.argv[1:]:
    os.environ["CFB_GROUP"] = g
    ctx = capi.Context(ix)
    db = ctx.upload(batch)
    for _ in range(2): ctx.classify_resident(db)
    ms = np.zeros(5)
    for _ in range(3): ms += np.array(ctx.classify_resident(db)[0])
    ms /= 3
    off, rec = ctx.resident_result()
    if ref is None: ref = (off, rec)
    same = np.array_equal(off, ref[0]) and np.array_equal(rec, ref[1])
    print("G=%s search %.2f ms prep %.2f resolve %.2f score %.2f total %.2f ms -> %.1f M reads/s  same_as_first=%s" % (g, ms[0], ms[1], ms[2], ms[3], ms[4], n / ms[4] / 1e3, same), flush=True)
    ctx.close()
