#!/usr/bin/env python3
"""File-to-file throughput of the drop-in `centrifuge-class` (FASTQ in, TSV + report out) on a GPU box:
the text operator (device tokeniser/formatter) against the record-level host reader and the reference
binary on the same files; with CFB_CLI_DEVICES (e.g. 0-7) also `--devices`, whose output must equal the one-GPU bytes.
Env: CFB_CLI_GBP (default 1 -> 1 Gbp index), CFB_CLI_READS (10M), CFB_CLI_RDLEN (100)."""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

EXE = os.path.join(ROOT, "centrifuge_b200", "centrifuge-class")
REF = os.path.join(ROOT, "oracle", "_ref", "centrifuge-class")


def run(exe, args, env=None):
    t0 = time.time()
    p = subprocess.run([exe] + args, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, **(env or {})))
    dt = time.time() - t0
    if p.returncode != 0:
        raise RuntimeError(p.stderr.decode()[-2000:])
    return dt, p.stderr.decode()


def main():
    n = int(os.environ.get("CFB_CLI_READS", 10000000)); L = int(os.environ.get("CFB_CLI_RDLEN", 100))
    sys.argv = ["bench.py", "--index-gbp", os.environ.get("CFB_CLI_GBP", "1"), "--rdlen", str(L)]
    a = bench.parse_args()
    base, d = bench.get_index(a)
    work = os.environ.get("CFB_CLI_DIR", "/dev/shm" if os.path.isdir("/dev/shm") else d)
    fq = os.path.join(work, "cli_bench.fq"); small = os.path.join(work, "cli_bench_small.fq")
    t0 = time.time()
    with open(fq, "wb") as f:
        for s in range(0, n, 2000000):
            k = min(2000000, n - s)
            m = bench.make_reads(a, k, 1000 + s).fastq(0, start=s)
            if s == 0:
                with open(small, "wb") as g:
                    g.write(bench.make_reads(a, 200000, 1000).fastq(0).tobytes())
            f.write(m.tobytes())
    print("[cli_bench] %d reads, %.2f GB FASTQ written in %.1f s" % (n, os.path.getsize(fq) / 1e9, time.time() - t0), flush=True)
    out = os.path.join(work, "cli_out.tsv"); rep = os.path.join(work, "cli_out.rep")

    def stats(err):
        return " | ".join(l for l in err.splitlines() if l.startswith("[cfb]"))

    runs = [("text operator", [], {}), ("text operator (2nd run)", [], {})]
    for t in os.environ.get("CFB_CLI_THREAD_SWEEP", "").split(","):
        if t:
            runs.append(("text operator, %s read threads" % t, [], {"CFB_READ_THREADS": t}))
    runs.append(("host reader", ["--host-parse", "-u", str(min(n, 2000000))], {}))
    devs = os.environ.get("CFB_CLI_DEVICES")
    if devs:
        runs.append(("text operator, --devices %s" % devs, ["--devices", devs], {}))
    for tag, extra, env in runs:
        o_, r_ = (out + ".dev", rep + ".dev") if "--devices" in extra else (out, rep)
        dt, err = run(EXE, ["-q", "-x", base, "-U", fq, "-S", o_, "--report-file", r_] + extra, dict(env, CFB_TEXT_STATS="1"))
        print("[cli_bench] %s: wall %.2f s; %s" % (tag, dt, stats(err)), flush=True)
        if "--devices" in extra:
            a_ = subprocess.run(["cmp", "-s", o_, out]).returncode == 0 and subprocess.run(["cmp", "-s", r_, rep]).returncode == 0
            print("[cli_bench] --devices output identical to the one-GPU output (TSV + report): %s" % a_, flush=True)
    # same bytes as the reference on a subset
    dt, err = run(EXE, ["-q", "-x", base, "-U", small, "-S", out + ".s", "--report-file", rep + ".s"], {"CFB_TEXT_STATS": "1"})
    if os.path.exists(REF):
        for p in (1, 16):
            t, _ = run(REF, ["-q", "-p", str(p), "-x", base, "-U", small, "-S", out + ".r", "--report-file", rep + ".r"])
            print("[cli_bench] reference -p %d on 200000 reads: wall %.2f s (incl. index load)" % (p, t), flush=True)
            if p == 1:
                same = open(out + ".s", "rb").read() == open(out + ".r", "rb").read() and open(rep + ".s", "rb").read() == open(rep + ".r", "rb").read()
                print("[cli_bench] TSV + report identical to the reference on the subset: %s" % same, flush=True)
    for f in (fq, small, out, rep, out + ".s", rep + ".s", out + ".r", rep + ".r", out + ".dev", rep + ".dev"):
        if os.path.exists(f):
            os.remove(f)


if __name__ == "__main__":
    main()
