#!/usr/bin/env python3
"""Reference binary throughput vs -p on the bench index (slope between two sample sizes excludes index load)."""
import os, sys, time, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
g = int(os.environ.get("CFB_BENCH_GENERA", 900))
base, d = bench.get_index(g, 10, 1000000, 12345)
n1, n2 = 20000, 420000
for n in (n1, n2):
    fq = os.path.join(d, "sample_%d.fq" % n)
    if not os.path.exists(fq):
        bench.write_fastq(fq, bench.make_reads(g, 10, 1000000, 12345, n, 100, 999))
for p in [int(x) for x in sys.argv[1:]]:
    t1 = bench.ref_reads_per_s(base, os.path.join(d, "sample_%d.fq" % n1), p)
    t2 = bench.ref_reads_per_s(base, os.path.join(d, "sample_%d.fq" % n2), p)
    print("-p %d: %d reads %.2f s, %d reads %.2f s -> %.0f reads/s excluding load" % (p, n1, t1, n2, t2, (n2 - n1) / (t2 - t1)), flush=True)
