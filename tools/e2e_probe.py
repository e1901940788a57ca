#!/usr/bin/env python3
"""Where does the e2e (host buffers in / records out) time go?  per-call submit / wait latencies."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from centrifuge_b200 import capi
genera = int(os.environ.get("CFB_BENCH_GENERA", 900)); n = int(os.environ.get("CFB_BENCH_READS", 2000000))
base, d = bench.get_index(genera, 10, 1000000, 12345)
codes = bench.make_reads(genera, 10, 1000000, 12345, n, 100, 1000)
bases = capi.pinned_array((n * 100,), np.uint8); bases[:] = codes.reshape(-1)
lens = capi.pinned_array((n,), np.uint32); lens[:] = 100
offs = capi.pinned_array((n,), np.uint64); offs[:] = np.arange(n, dtype=np.uint64) * np.uint64(100)
flags = capi.pinned_array((n,), np.uint8); flags[:] = ((codes == 4).sum(axis=1) <= 15).astype(np.uint8)
batch = capi.make_batch(bases, offs, lens, None, None, flags)
ix = capi.Index(base, 0); ctx = capi.Context(ix)
for _ in range(3):
    ctx.submit(0, batch); ctx.wait(0, copy=False)
for depth in (1, 2, 3, 4):
    ts, tw = [], []
    t0 = time.perf_counter(); infl = []
    K = 10
    for s in range(K):
        if len(infl) == depth:
            a = time.perf_counter(); ctx.wait(infl.pop(0), copy=False); tw.append(time.perf_counter() - a)
        a = time.perf_counter(); ctx.submit(s % depth, batch); ts.append(time.perf_counter() - a); infl.append(s % depth)
    while infl:
        a = time.perf_counter(); ctx.wait(infl.pop(0), copy=False); tw.append(time.perf_counter() - a)
    tot = time.perf_counter() - t0
    print("depth %d: %.2f ms/step (%.1f M reads/s); submit avg %.2f ms, wait avg %.2f ms" % (depth, 1000 * tot / K, n * K / tot / 1e6, 1000 * np.mean(ts), 1000 * np.mean(tw)))
