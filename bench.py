#!/usr/bin/env python3
"""bench.py -- reads/s of the classification hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic 100 bp single-end reads against a
synthetic genus/species index (SURVEY.md Appendix C recipe) replicated in each GPU's HBM.

  value     reads/s with the batch already resident in HBM (kernels only, CUDA events, max over ranks)
  e2e       reads/s through cfb_classify_submit/wait with HOST buffers: every step's H2D of the reads and D2H of the
            result records are inside the timed region; a step is cut into sub-batches on half of the context's
            slots and steps stream (step k+1 is submitted before step k's records are collected)
  e2e_text  the same from FASTQ bytes to TSV bytes through the text-level operator (cfb_text_submit/wait)
  roofline  k_search: algorithmic bytes (128 B per side touched + 16 B per ftab probe, counted by the
            kernel's own counters in a separate un-timed pass) / its CUDA-event time, vs measured HBM peak
  cpu_baseline   the unmodified reference binary (oracle/_ref/centrifuge-class -p <cores>) on a bounded sample

`--impl reference` times that reference binary instead (all host threads, bounded sample per step).
Multi-GPU: one process per GPU (torchrun), reads sharded, index replicated, one NCCL all-reduce of the
dense per-taxon count vector at the end of every step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

REF_CLASS = os.path.join(ROOT, "oracle", "_ref", "centrifuge-class")
REF_BUILD = os.path.join(ROOT, "oracle", "_ref", "centrifuge-build-bin")
CACHE = os.environ.get("CFB_BENCH_CACHE", "/tmp/cfb200_bench")


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


# ------------------------------------------------------------------------------------ workload
PREFIX = "cid"      # sequence names cid<i>: >= 10 of them make the index a "compressed" one (ihits = 20, bt2_idx.h:648-663), as p_compressed is


def synth_opts(genera, species, length, seed, base=None, device=0, tax=None):
    from centrifuge_b200 import capi
    kw = {}
    if tax:
        kw = dict(conversion_table=tax[0], taxonomy_tree=tax[1], name_table=tax[2])
    return capi.build_opts(base, synth=(genera, species, length, seed, 0.03), device=device, synth_prefix=PREFIX, **kw)


def get_index(genera, species, length, seed, device=0):
    """Synthetic p_compressed-class index on local disk, built once per box by the GPU builder
    (centrifuge_b200/csrc/cf_build.cu; byte-identical to centrifuge-build-bin, tests/test_gpu_build.py)."""
    from centrifuge_b200 import capi
    tag = "%s_g%d_s%d_l%d_seed%d" % (PREFIX, genera, species, length, seed)
    d = os.path.join(CACHE, tag)
    base = os.path.join(d, "idx")
    if os.path.exists(os.path.join(d, "done")):
        return base, d
    os.makedirs(d, exist_ok=True)
    t0 = time.time()
    tax = capi.write_synth_taxonomy(d, genera, species, length, prefix=PREFIX)
    capi.build_index(synth_opts(genera, species, length, seed, base, device, tax))
    open(os.path.join(d, "done"), "w").close()
    log("index %s built on the GPU in %.1f s" % (tag, time.time() - t0))
    return base, d


def make_reads(genera, species, length, gseed, n, rdlen, seed, device=0):
    """cfb_synth_reads: uniform genome/position/strand, 1% substitutions, 0.1% N, 5% random reads.
    Returns codes (n, rdlen) uint8 in 0..4."""
    from centrifuge_b200 import capi
    return capi.synth_reads(synth_opts(genera, species, length, gseed, device=device), n, rdlen, seed)


def fastq_matrix(codes, start=0):
    """Fixed-width FASTQ records ("@r%09d", 4 lines) as one uint8 matrix, vectorised."""
    n, L = codes.shape
    m = np.empty((n, 2 + 9 + 1 + L + 3 + L + 1), dtype=np.uint8)
    m[:, 0] = ord("@"); m[:, 1] = ord("r")
    idx = np.arange(start, start + n, dtype=np.int64)
    m[:, 2:11] = (idx[:, None] // (10 ** np.arange(8, -1, -1, dtype=np.int64))[None, :] % 10 + 48).astype(np.uint8)
    m[:, 11] = 10
    m[:, 12:12 + L] = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes]
    m[:, 12 + L] = 10; m[:, 13 + L] = ord("+"); m[:, 14 + L] = 10
    m[:, 15 + L:15 + 2 * L] = ord("I")
    m[:, 15 + 2 * L] = 10
    return m


def write_fastq(path, codes, prefix="r"):
    asc = np.frombuffer(b"ACGTN", dtype=np.uint8)[codes]
    n, L = asc.shape
    q = b"I" * L
    with open(path, "wb") as f:
        for i in range(n):
            f.write(b"@%s%d\n" % (prefix.encode(), i) + asc[i].tobytes() + b"\n+\n" + q + b"\n")


def bind_to_gpu_numa_node(dev):
    """Run this rank on the CPUs of the NUMA node its GPU hangs off, so that the pinned host buffers it allocates
    (first touch) and the threads that fill them are local to that GPU's PCIe root.  Best effort."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(dev)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]                                   # sysfs uses 4-digit domains
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001
        pass
    return None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed regions, via NVML in-process (no nvidia-smi forks, which
    contend for the driver lock with the CUDA calls being timed)."""

    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.samples, self.reasons, self.stop_flag, self.maxmhz = dev, [], set(), False, None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.dev)
            self.maxmhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0)),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0)),
                    "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0))}
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self.stop_flag:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(h))
                for nm, b in bits.items():
                    if b and (r & b):
                        self.reasons.add(nm)
                time.sleep(0.05)
        except Exception as e:  # noqa: BLE001
            self.reasons.add("nvml_unavailable: %s" % e)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.maxmhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def ref_reads_per_s(base, fq, threads):
    t0 = time.time()
    subprocess.check_call([REF_CLASS, "-q", "-x", base, "-U", fq, "-p", str(threads), "-S", "/dev/null", "--report-file", "/dev/null"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.time() - t0


class RefArm:
    """The unmodified reference binary on the host cores.  Index load is excluded by differencing two
    sample sizes; the thread count is the best of a short sweep (the reference's one-read-per-mutex
    input loop, pat.h:786-811, stops scaling well before 128 threads)."""

    def __init__(self, a, base, d, device=0):
        self.base, self.d = base, d
        self.n_small, self.n = 20000, a.cpu_sample
        self.fq_small, self.fq = os.path.join(d, "sample_%d.fq" % self.n_small), os.path.join(d, "sample_%d.fq" % self.n)
        for n, fq in ((self.n_small, self.fq_small), (self.n, self.fq)):
            if not os.path.exists(fq):
                write_fastq(fq, make_reads(a.genera, a.species, a.genome_len, 12345, n, a.rdlen, 999, device))
        ncores = os.cpu_count() or 1
        self.load_s = {}
        best = None
        for p in sorted(set(min(x, ncores) for x in (8, 16, 24, 32, 64, ncores))):
            t1 = ref_reads_per_s(base, self.fq_small, p)
            t2 = ref_reads_per_s(base, self.fq, p)
            rate = (self.n - self.n_small) / max(t2 - t1, 1e-6)
            self.load_s[p] = t1
            log("reference -p %d: %.0f reads/s (%.1f s for %d reads, %.1f s for %d)" % (p, rate, t2, self.n, t1, self.n_small))
            if best is None or rate > best[1]:
                best = (p, rate)
            elif rate < 0.5 * best[1]:
                break
        self.threads, self.sweep_rate = best

    def step(self):
        """seconds of classification work for self.n - self.n_small reads (index load differenced out)"""
        t2 = ref_reads_per_s(self.base, self.fq, self.threads)
        t1 = ref_reads_per_s(self.base, self.fq_small, self.threads)
        return max(t2 - t1, 1e-6), self.n - self.n_small


def parity_check(a, ctx, ix, base, d, codes, nsample):
    """Un-timed: the first `nsample` reads of this run's own batch, classified by the *same* context the timed loops
    use (every derived table live), against the unmodified reference binary on the same index and the same FASTQ:
    (1) the TSV the text operator returns must equal the reference's bytes, (2) the records the C ABI returns must
    be the rows of that TSV (per read: the records with the best score <-> the rows, taxID / score / hitLength)."""
    from centrifuge_b200 import capi
    import pandas as pd
    tb = ix.tables()
    tables = {"ftabk": tb["ftabk_chars"], "rtab": 8 * tb["resolve_entry_bytes"], "walk8": tb["walk8_bytes"] > 0,
              "compressed": bool(ix.info.compressed), "rows_beyond_2^32": bool(ix.info.len >= (1 << 32))}
    if not os.path.exists(REF_CLASS):
        return {"reads": 0, "identical": None, "tables": tables, "skipped": "oracle/_ref/centrifuge-class not shipped"}
    sub = codes[:nsample]
    n, L = sub.shape
    fm = fastq_matrix(sub)
    fq, ref_tsv = os.path.join(d, "parity_%d.fq" % os.getpid()), os.path.join(d, "parity_%d.tsv" % os.getpid())
    with open(fq, "wb") as f:
        f.write(fm.tobytes())
    p = min(16, os.cpu_count() or 1)
    t0 = time.time()
    subprocess.check_call([REF_CLASS, "-q", "-x", base, "-U", fq, "-p", str(p), "--reorder", "-S", ref_tsv, "--report-file", "/dev/null"],
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    with open(ref_tsv, "rb") as f:
        want = f.read()
    t_ref = time.time() - t0
    pt = capi.pinned_array((fm.size,), np.uint8); pt[:] = fm.reshape(-1)
    ctx.text_submit(0, pt, None, n, maxlen_hint=L)
    r = ctx.text_wait(0, discard=True)
    header = b"readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches\n"
    tsv_ok = (not r["irregular"]) and header + r["tsv"] == want
    # records of the C ABI against the reference's rows
    flags = ((sub == 4).sum(axis=1) <= int(0.15 * L)).astype(np.uint8)
    lens = np.full(n, L, dtype=np.uint32); offs = np.arange(n, dtype=np.uint64) * np.uint64(L)
    off, recs = ctx.classify(capi.make_batch(np.ascontiguousarray(sub.reshape(-1)), offs, lens, None, None, flags))
    df = pd.read_csv(ref_tsv, sep="\t", dtype={"readID": str, "seqID": str})
    rid = df["readID"].str[1:].astype(np.int64).to_numpy()
    cls = (df["seqID"] != "unclassified").to_numpy()
    rows = np.stack([rid[cls], df["taxID"].to_numpy(np.int64)[cls], df["score"].to_numpy(np.int64)[cls], df["hitLength"].to_numpy(np.int64)[cls]], axis=1)
    cnt = np.diff(off.astype(np.int64))
    unit = np.repeat(np.arange(n), cnt)
    best = np.zeros(n, dtype=np.int64)
    np.maximum.at(best, unit, recs["score"].astype(np.int64))
    top = recs["score"].astype(np.int64) == best[unit]
    mine = np.stack([unit[top], recs["taxid"][top].astype(np.int64), recs["score"][top].astype(np.int64), recs["hitlen"][top].astype(np.int64)], axis=1)
    key = lambda x: x[np.lexsort((x[:, 3], x[:, 2], x[:, 1], x[:, 0]))]
    rec_ok = mine.shape == rows.shape and bool(np.array_equal(key(mine), key(rows)))
    uncl_ok = bool(np.array_equal(np.sort(rid[~cls]), np.nonzero(cnt == 0)[0]))
    os.unlink(fq); os.unlink(ref_tsv)
    return {"reads": int(n), "identical": bool(tsv_ok and rec_ok and uncl_ok), "tsv_bytes_identical": bool(tsv_ok), "abi_records_match_rows": bool(rec_ok and uncl_ok),
            "tables": tables, "rows": int(len(df)), "reference": "oracle/_ref/centrifuge-class -p %d --reorder, same index, same FASTQ (%.1f s)" % (p, t_ref)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------ main
def main():
    # stdout carries exactly one JSON line: route everything else that native libraries print to fd 1
    # (e.g. NCCL's version banner) to stderr, and keep a private handle for the result line
    sys.stdout.flush()
    result_fd = os.dup(1)
    os.dup2(2, 1)
    result = os.fdopen(result_fd, "w")
    try:
        return _main(result)
    finally:
        result.flush()


def _main(result):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cfb200", choices=["cfb200", "reference"])
    ap.add_argument("--genera", type=int, default=int(os.environ.get("CFB_BENCH_GENERA", 900)))
    ap.add_argument("--species", type=int, default=int(os.environ.get("CFB_BENCH_SPECIES", 10)))
    ap.add_argument("--genome-len", type=int, default=int(os.environ.get("CFB_BENCH_GENOME_LEN", 1000000)))
    ap.add_argument("--reads", type=int, default=int(os.environ.get("CFB_BENCH_READS", 2000000)), help="reads per step per GPU")
    ap.add_argument("--rdlen", type=int, default=100)
    ap.add_argument("--cpu-sample", type=int, default=int(os.environ.get("CFB_BENCH_CPU_SAMPLE", 400000)))
    ap.add_argument("--parity-reads", type=int, default=int(os.environ.get("CFB_BENCH_PARITY_READS", 200000)), help="reads of the step batch checked against the reference binary (un-timed)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    ncores = os.cpu_count() or 1
    workload = "synthetic %d genera x %d species x %d bp index (%.0f Mbp), %d x %d bp SE reads per step per GPU" % (
        a.genera, a.species, a.genome_len, a.genera * a.species * a.genome_len / 1e6, a.reads, a.rdlen)

    if a.impl == "reference":
        if rank != 0:
            return 0
        base, d = get_index(a.genera, a.species, a.genome_len, 12345)
        arm = RefArm(a, base, d)
        for _ in range(min(a.warmup, 1)):
            arm.step()
        tot, nreads = 0.0, 0
        for _ in range(a.steps):
            t, n = arm.step()
            tot += t; nreads += n
        val = nreads / tot
        sample = "%d reads per step (bounded sample), centrifuge-class -p %d (best of a thread sweep up to %d), FASTQ in, TSV to /dev/null, index load differenced out" % (
            nreads // a.steps, arm.threads, ncores)
        print(json.dumps({"metric": "reads/sec (%d bp SE classification)" % a.rdlen, "value": val, "unit": "reads/s", "n_gpus": a.gpus, "steps": a.steps,
                          "warmup": min(a.warmup, 1), "ms_per_step": 1000 * tot / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "u64", "data": "synthetic", "impl": "reference",
                          "config": {"workload": workload, "sample": sample},
                          "cpu_baseline": {"value": val, "unit": "reads/s", "cores": arm.threads, "kind": "reference", "sample": sample},
                          "e2e": {"value": val, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), file=result, flush=True)
        return 0

    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    import torch
    from centrifuge_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the cfb200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        base, d = get_index(a.genera, a.species, a.genome_len, 12345, local)
    if dist:
        dist.barrier()
    base, d = get_index(a.genera, a.species, a.genome_len, 12345, local)
    codes = make_reads(a.genera, a.species, a.genome_len, 12345, a.reads, a.rdlen, 1000 + rank, local)
    n = a.reads
    # host buffers of the e2e arm live in pinned memory (cfb_host_alloc), as a real caller's parse buffers would
    bases = capi.pinned_array((n * a.rdlen,), np.uint8); bases[:] = codes.reshape(-1)
    lens = capi.pinned_array((n,), np.uint32); lens[:] = a.rdlen
    offs = capi.pinned_array((n,), np.uint64); offs[:] = np.arange(n, dtype=np.uint64) * np.uint64(a.rdlen)
    flags = capi.pinned_array((n,), np.uint8); flags[:] = ((codes == 4).sum(axis=1) <= int(0.15 * a.rdlen)).astype(np.uint8)
    batch = capi.make_batch(bases, offs, lens, None, None, flags)

    t0 = time.time()
    ix = capi.Index(base, local)
    log("rank %d: index in HBM: %.2f GB in %.1f s (%d sides)" % (rank, ix.info.device_bytes / 1e9, time.time() - t0, ix.info.num_sides))
    ctx = capi.Context(ix)
    from centrifuge_b200.abundance import taxon_counts
    node_taxids = ix.node_taxids()
    # dense per-taxon {numReads, numUniqueReads} of this rank's shard: folded once from the first
    # result on the host, then all-reduced (NCCL, sum) at the end of every step
    off0, recs0 = ctx.classify(batch)
    local_counts = taxon_counts(node_taxids, off0, recs0, k=5)
    counts = torch.from_numpy(local_counts.reshape(-1).copy()).cuda()

    def sync_all():
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # ---------------- un-timed parity check of the configuration being timed (rank 0; the other ranks wait at the barrier)
    parity = None
    if rank == 0:
        parity = parity_check(a, ctx, ix, base, d, codes, min(a.parity_reads, n))
        log("parity check: %s" % json.dumps(parity))

    # ---------------- un-timed counter pass (algorithmic bytes of this exact batch)
    os.environ["CFB_COUNT"] = "1"
    cctx = capi.Context(ix)
    del os.environ["CFB_COUNT"]
    cctx.classify(batch)
    ctr = cctx.counters()
    cctx.close()
    sample_w = ix.info.sample_bytes
    bytes_search = 128 * ctr["sides_search"] + 16 * ctr["ftab_probes"]
    bytes_walk = 128 * ctr["walk_steps"] + sample_w * ctr["rows_resolved"]

    # ---------------- value: resident batch, kernels only
    dbatch = ctx.upload(batch)
    for _ in range(a.warmup):
        ctx.classify_resident(dbatch)
    sampler = ClockSampler(local); sampler.start()
    sync_all()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    kms = np.zeros(5)
    for _ in range(a.steps):
        ms, nrec = ctx.classify_resident(dbatch)       # per-stage CUDA-event times on the kernels' own stream
        kms += np.array(ms)
        if dist:
            step_counts = counts.clone()
            dist.all_reduce(step_counts)
    sync_all()
    wall = time.perf_counter() - t_wall0
    launches_value = ctx.launches()
    # device time of the timed region = sum of the per-step event spans (each step is synchronised)
    step_ms = kms[4] / a.steps
    t_dev = torch.tensor([kms[4] / 1000.0, wall], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_s, wall_s = float(t_dev[0]), float(t_dev[1])
    value = world * n * a.steps / dev_s

    # ---------------- e2e: host buffers in, host records out.  One step = the same batch, cut into
    # n_slots sub-batches that are submitted back to back (H2D of one overlaps the kernels of another)
    # and then waited for: every step is self-contained, all copies are inside the timed region.
    # Steps stream: the sub-batches of step k+1 are submitted (to the other half of the slots) before the results of
    # step k are collected, as a caller with a steady supply of reads would do; every step's copies are inside the
    # timed region, which ends when the last step's records are in host memory.
    nslots = max(1, ctx.n_slots // 2)
    m = n // nslots
    sub = []
    offs_sub = capi.pinned_array((m,), np.uint64); offs_sub[:] = np.arange(m, dtype=np.uint64) * np.uint64(a.rdlen)
    for sl in range(nslots):
        sub.append(capi.make_batch(bases[sl * m * a.rdlen:(sl + 1) * m * a.rdlen], offs_sub, lens[sl * m:(sl + 1) * m], None, None, flags[sl * m:(sl + 1) * m]))
    n_e2e = m * nslots

    def e2e_submit(step):
        for sl in range(nslots):
            ctx.submit((step % 2) * nslots + sl, sub[sl])

    def e2e_collect(step):
        nr = 0
        for sl in range(nslots):
            nr += ctx.wait((step % 2) * nslots + sl, copy=False)[1]
        return nr

    def e2e_run(steps):
        nr = 0
        e2e_submit(0)
        for s in range(1, steps):
            e2e_submit(s)
            nr += e2e_collect(s - 1)
            if dist:
                step_counts = counts.clone()
                dist.all_reduce(step_counts)
        nr += e2e_collect(steps - 1)
        if dist:
            step_counts = counts.clone()
            dist.all_reduce(step_counts)
        return nr

    e2e_run(max(2, min(a.warmup, 2)))                  # every slot allocates its buffers before the timed region
    sync_all()
    t0 = time.perf_counter()
    d2h = e2e_run(a.steps) * 24 + a.steps * (n_e2e + nslots) * 4
    sync_all()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * n_e2e * a.steps / float(te[0])

    # ---------------- e2e_text: FASTQ bytes in (pinned host memory), classification TSV bytes out, through the
    # text-level operator (device tokeniser / selector / formatter): what the CLI does minus the file system.
    txt = []
    for sl in range(nslots):
        fm = fastq_matrix(codes[sl * m:(sl + 1) * m], sl * m)
        pt = capi.pinned_array((fm.size,), np.uint8); pt[:] = fm.reshape(-1)
        txt.append(pt)

    def text_submit(step):
        for sl in range(nslots):
            ctx.text_submit((step % 2) * nslots + sl, txt[sl], None, m, maxlen_hint=a.rdlen)

    def text_collect(step):
        nb = 0
        for sl in range(nslots):
            r = ctx.text_wait((step % 2) * nslots + sl, copy=False)
            if r["irregular"]:
                raise RuntimeError("text operator rejected the synthetic FASTQ")
            nb += r["tsv_bytes"] + r["n_multi"] * 8 * 6
        return nb

    def text_run(steps):
        nb = 0
        text_submit(0)
        for s in range(1, steps):
            text_submit(s)
            nb += text_collect(s - 1)
        return nb + text_collect(steps - 1)

    text_run(max(2, min(a.warmup, 2)))
    sync_all()
    t0 = time.perf_counter()
    tsv_bytes = text_run(a.steps)
    sync_all()
    tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_text = world * n_e2e * a.steps / float(tt[0])
    sampler.stop_flag = True; sampler.join(timeout=2)
    launches_total = ctx.launches()
    h2d = bases.nbytes + offs.nbytes + lens.nbytes + flags.nbytes

    peak, peak_src = measured_peak()
    search_s = kms[0] / 1000.0 / a.steps
    achieved = bytes_search / search_s / 1e9
    # measured DRAM traffic of the kernel (one ncu --set full capture, profiles/r01_traffic.json), scaled to this launch
    traffic, gather = None, None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_read"] * n
        sect = tj["dram_bytes_read"] / 32.0 / tj["reads_in_capture"] * n          # 32-byte DRAM sectors read per launch
        gather = {"what": "random 32-byte sector gathers: achieved vs the ceiling measured on this part by tools/gather_bench.cu (~34.5 G sectors/s)",
                  "achieved_gsectors_s": sect / search_s / 1e9, "ceiling_gsectors_s": 34.5, "frac": sect / search_s / 1e9 / 34.5}
    out = {
        "metric": "reads/sec (%d bp SE classification)" % a.rdlen, "value": value, "unit": "reads/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1000 * dev_s / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload, "index_bytes_hbm": int(ix.info.device_bytes), "l2": "index replica %.0f MB vs 126 MB L2; same batch re-walked every step" % (ix.info.device_bytes / 1e6),
                   "parallelism": "reads sharded over %d GPU(s), index replicated, 1 NCCL all-reduce of per-taxon counts per step" % world},
        "e2e": {"value": e2e, "unit": "reads/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h // max(a.steps, 1))},
        "e2e_text": {"value": e2e_text, "unit": "reads/s", "what": "FASTQ text in pinned host memory -> TSV rows in host memory (cfb_text_submit/wait)",
                     "h2d_bytes_per_step": int(sum(t.nbytes for t in txt)), "d2h_bytes_per_step": int(tsv_bytes // max(a.steps, 1))},
        "gpu_launches": int(launches_value),
        "roofline": {"bound": "hbm", "kernel": "k_search_t", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "random_gather": gather,
                     "note": "achieved counts the reference algorithm's bytes (SURVEY 8d); jump tables (K-mer table, walk8) skip part of them, so frac can exceed 1: the physical bound is random_gather",
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": int(bytes_search), "kernel_ms": 1000 * search_s,
                     "sides_per_read": ctr["sides_search"] / max(ctr["units"], 1), "walk_bytes_per_launch": int(bytes_walk)},
        "kernel_ms": {"search": kms[0] / a.steps, "prep_rows": kms[1] / a.steps, "resolve": kms[2] / a.steps, "score_compact": kms[3] / a.steps, "total": step_ms},
        "clocks": sampler.summary(),
        "taxon_vector": {"len": int(counts.numel()), "classified_reads_rank0": int(local_counts[:-1, 0].sum()), "unclassified_reads_rank0": int(local_counts[-1, 1])},
        "wall_s_value_region": wall_s,
        "parity_check": parity,
    }
    out["config"]["tables"] = {k: v for k, v in ix.tables().items()}
    if numa is not None:
        out["config"]["numa_node"] = numa
    if rank == 0 and world > 1:
        out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": 0, "kind": "reference", "sample": "reported by the N=1 run only"}
        print(json.dumps(out), file=result, flush=True)
    elif rank == 0:
        # bounded CPU baseline: the unmodified reference binary on the host cores
        try:
            arm = RefArm(a, base, d, local)
            tt, nn = arm.step()
            out["cpu_baseline"] = {"value": nn / tt, "unit": "reads/s", "cores": arm.threads, "kind": "reference",
                                   "sample": "%d reads, centrifuge-class -p %d (best of a sweep up to %d threads), FASTQ in, TSV to /dev/null, index load differenced out" % (nn, arm.threads, ncores)}
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": ncores, "kind": "reference", "sample": "failed: %s" % e}
        print(json.dumps(out), file=result, flush=True)
    ctx.close(); ix.close()
    if dist:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and parity.get("identical") is False:
        log("PARITY CHECK FAILED: the timed configuration does not reproduce the reference's output")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
