#!/usr/bin/env python3
"""Seeded synthetic genomes / taxonomy / reads for parity tests and benchmarks.

Recipe follows SURVEY.md Appendix C / section 8(d): G genera x S species; a genus
"base" is iid uniform ACGT of length L, every other species of the genus copies the
base with per-base substitution probability `div` (default 3%), which creates
multi-genome partial hits and exercises the taxonomy-tree reduction of the
classifier.  Taxonomy: 1 (root, no rank) -> 100+g (genus) -> 1000+s (species)
[-> 100000+i (no rank, one per sequence) with --strains].

Nothing here is used by the product path; it only produces inputs.
"""
import argparse
import os
import sys

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
COMP = {ord("A"): "T", ord("C"): "G", ord("G"): "C", ord("T"): "A", ord("N"): "N"}


def make_genomes(genera, species, length, seed, div=0.03):
    """Return list of uint8 arrays (values 0..3), one per sequence."""
    rng = np.random.default_rng(seed)
    seqs = []
    for g in range(genera):
        base = rng.integers(0, 4, size=length, dtype=np.uint8)
        for s in range(species):
            if s == 0:
                seqs.append(base)
                continue
            cp = base.copy()
            mask = rng.random(length) < div
            n = int(mask.sum())
            cp[mask] = (cp[mask] + rng.integers(1, 4, size=n, dtype=np.uint8)) & 3
            seqs.append(cp)
    return seqs


def write_genomes(outdir, genera, species, length, seed, div=0.03, cid=False,
                  strains=False, width=80):
    os.makedirs(outdir, exist_ok=True)
    seqs = make_genomes(genera, species, length, seed, div)
    prefix = "cid" if cid else "seq"
    with open(os.path.join(outdir, "genomes.fa"), "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">%s%d\n" % (prefix.encode(), i))
            a = ACGT[s]
            full = (len(a) // width) * width
            if full:
                body = a[:full].reshape(-1, width)
                nl = np.full((body.shape[0], 1), 10, dtype=np.uint8)
                f.write(np.hstack([body, nl]).tobytes())
            if full < len(a):
                f.write(a[full:].tobytes() + b"\n")
    with open(os.path.join(outdir, "conv.tsv"), "w") as f:
        for i in range(len(seqs)):
            tid = (100000 + i) if strains else (1000 + i)
            f.write("%s%d\t%d\n" % (prefix, i, tid))
    with open(os.path.join(outdir, "nodes.dmp"), "w") as f:
        f.write("1\t|\t1\t|\tno rank\t|\n")
        for g in range(genera):
            f.write("%d\t|\t1\t|\tgenus\t|\n" % (100 + g))
        for i in range(len(seqs)):
            f.write("%d\t|\t%d\t|\tspecies\t|\n" % (1000 + i, 100 + i // species))
            if strains:
                f.write("%d\t|\t%d\t|\tno rank\t|\n" % (100000 + i, 1000 + i))
    with open(os.path.join(outdir, "names.dmp"), "w") as f:
        f.write("1\t|\troot\t|\t\t|\tscientific name\t|\n")
        for g in range(genera):
            f.write("%d\t|\tGenus%d\t|\t\t|\tscientific name\t|\n" % (100 + g, g))
        for i in range(len(seqs)):
            f.write("%d\t|\tGenus%d species%d\t|\t\t|\tscientific name\t|\n"
                    % (1000 + i, i // species, i))
            if strains:
                f.write("%d\t|\tGenus%d species%d strain%d\t|\t\t|\tscientific name\t|\n"
                        % (100000 + i, i // species, i, i))
    return seqs


def sample_reads(seqs, n, rdlen, seed, sub=0.01, nrate=0.001, random_frac=0.05,
                 lens=None):
    """Yield (name, uint8 ascii sequence) tuples.  lens=(lo,hi) -> uniform lengths."""
    rng = np.random.default_rng(seed)
    out = []
    nseq = len(seqs)
    for k in range(n):
        L = rdlen if lens is None else int(rng.integers(lens[0], lens[1] + 1))
        if rng.random() < random_frac:
            r = rng.integers(0, 4, size=L, dtype=np.uint8)
            name = "r%d_rand" % k
        else:
            si = int(rng.integers(0, nseq))
            s = seqs[si]
            L = min(L, len(s))
            p = int(rng.integers(0, len(s) - L + 1))
            r = s[p:p + L].copy()
            m = rng.random(L) < sub
            r[m] = (r[m] + 1) & 3
            name = "r%d_s%d_p%d" % (k, si, p)
            if rng.random() < 0.5:
                r = (3 - r)[::-1]
        a = ACGT[r].copy()
        if nrate > 0:
            a[rng.random(L) < nrate] = ord("N")
        out.append((name, a))
    return out


def sample_pairs(seqs, n, rdlen, seed, sub=0.01, nrate=0.001, random_frac=0.05,
                 ins=(200, 500)):
    rng = np.random.default_rng(seed)
    out = []
    nseq = len(seqs)
    for k in range(n):
        if rng.random() < random_frac:
            r1 = rng.integers(0, 4, size=rdlen, dtype=np.uint8)
            r2 = rng.integers(0, 4, size=rdlen, dtype=np.uint8)
            name = "p%d_rand" % k
        else:
            si = int(rng.integers(0, nseq))
            s = seqs[si]
            frag = int(rng.integers(ins[0], ins[1] + 1))
            frag = min(max(frag, rdlen), len(s))
            p = int(rng.integers(0, len(s) - frag + 1))
            f = s[p:p + frag]
            r1 = f[:rdlen].copy()
            r2 = (3 - f[frag - rdlen:])[::-1].copy()
            for r in (r1, r2):
                m = rng.random(len(r)) < sub
                r[m] = (r[m] + 1) & 3
            if rng.random() < 0.5:
                r1, r2 = r2, r1
            name = "p%d_s%d_p%d" % (k, si, p)
        a1, a2 = ACGT[r1].copy(), ACGT[r2].copy()
        if nrate > 0:
            a1[rng.random(len(a1)) < nrate] = ord("N")
            a2[rng.random(len(a2)) < nrate] = ord("N")
        out.append((name, a1, a2))
    return out


def rc_codes(x):
    return (3 - x)[::-1]


def write_adversarial(outdir, seed=33, n_reads=3000):
    """Small index + reads that exercise the rare branches of the classifier:
    inverted repeats (both strands hit => extend / twin-removal), tandem and homopolymer
    repeats (SA ranges larger than ihits), a 300-copy dispersed repeat, `cid` sequence names
    (compressed-index ihits rule), N-rich / very short / IUPAC reads, names with spaces."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    seqs = []
    base = rng.integers(0, 4, size=60000, dtype=np.uint8)
    seqs.append(base)
    seqs.append(rc_codes(base))
    m = base.copy()
    mask = rng.random(len(m)) < 0.02
    m[mask] = (m[mask] + 1) & 3
    seqs.append(m)
    seqs.append(rc_codes(m[10000:40000]))
    unit = rng.integers(0, 4, size=7, dtype=np.uint8)
    seqs.append(np.concatenate([rng.integers(0, 4, size=5000, dtype=np.uint8), np.tile(unit, 800),
                                rng.integers(0, 4, size=5000, dtype=np.uint8)]))
    seqs.append(np.concatenate([np.zeros(3000, dtype=np.uint8), rng.integers(0, 4, size=3000, dtype=np.uint8),
                                np.full(3000, 3, dtype=np.uint8)]))
    rep = rng.integers(0, 4, size=150, dtype=np.uint8)
    parts = []
    for _ in range(300):
        parts.append(rng.integers(0, 4, size=200, dtype=np.uint8))
        parts.append(rep)
    seqs.append(np.concatenate(parts))
    rep2 = rng.integers(0, 4, size=120, dtype=np.uint8)
    parts = []
    for _ in range(30):
        parts.append(rng.integers(0, 4, size=300, dtype=np.uint8))
        parts.append(rep2)
    seqs.append(np.concatenate(parts))
    for _ in range(12):
        seqs.append(rng.integers(0, 4, size=4000, dtype=np.uint8))
    n = len(seqs)
    with open(os.path.join(outdir, "genomes.fa"), "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">cid%d\n" % i)
            f.write(ACGT[s].tobytes() + b"\n")
    with open(os.path.join(outdir, "conv.tsv"), "w") as f:
        for i in range(n):
            f.write("cid%d\t%d\n" % (i, 1000 + i))
    with open(os.path.join(outdir, "nodes.dmp"), "w") as f:
        f.write("1\t|\t1\t|\tno rank\t|\n")
        f.write("10\t|\t1\t|\tfamily\t|\n11\t|\t1\t|\tfamily\t|\n")
        for g in range(4):
            f.write("%d\t|\t%d\t|\tgenus\t|\n" % (100 + g, 10 + g % 2))
        for i in range(n):
            f.write("%d\t|\t%d\t|\tspecies\t|\n" % (1000 + i, 100 + i % 4))
    with open(os.path.join(outdir, "names.dmp"), "w") as f:
        f.write("1\t|\troot\t|\t\t|\tscientific name\t|\n")
        for t in (10, 11):
            f.write("%d\t|\tFam%d\t|\t\t|\tscientific name\t|\n" % (t, t))
        for g in range(4):
            f.write("%d\t|\tGen%d\t|\t\t|\tscientific name\t|\n" % (100 + g, g))
        for i in range(n):
            f.write("%d\t|\tSp %d\t|\t\t|\tscientific name\t|\n" % (1000 + i, i))
    reads = []
    for k in range(n_reads):
        si = int(rng.integers(0, n))
        s = seqs[si]
        L = min(int(rng.integers(20, 260)), len(s))
        p = int(rng.integers(0, len(s) - L + 1))
        r = s[p:p + L].copy()
        mm = rng.random(L) < 0.02
        r[mm] = (r[mm] + 1) & 3
        if rng.random() < 0.5:
            r = rc_codes(r)
        a = ACGT[r].copy()
        u = rng.random()
        if u < 0.15:
            a[rng.random(L) < 0.05] = ord("N")
        elif u < 0.2:
            a[rng.random(L) < 0.2] = ord("N")
        reads.append(("x%d" % k, a))
    s0 = seqs[0]
    for nm, a in (("short1", s0[:1]), ("short5", s0[:5]), ("short9", s0[:9]), ("short10", s0[:10]),
                  ("short21", s0[:21]), ("short22", s0[100:122]), ("short23", s0[100:123])):
        reads.append((nm, ACGT[a].copy()))
    reads.append(("allN", np.full(50, ord("N"), dtype=np.uint8)))
    reads.append(("polyA", np.full(100, ord("A"), dtype=np.uint8)))
    reads.append(("polyT", np.full(100, ord("T"), dtype=np.uint8)))
    reads.append(("iupac", np.frombuffer(b"ACGTRYKMACGTACGTNNACGTTTGGCCAAXBDHVACGTACGTAGCTAGCTAGCTAGCATCGATCGACTAGC", dtype=np.uint8).copy()))
    reads.append(("name with space/1", ACGT[seqs[2][500:600]].copy()))
    reads.append(("pal", np.concatenate([ACGT[s0[200:260]], ACGT[rc_codes(s0[200:260])]])))
    write_fasta(os.path.join(outdir, "reads.fa"), reads)
    write_fastq(os.path.join(outdir, "reads.fq"), reads, qual=b"5")
    return seqs, reads


def write_fasta(path, reads):
    with open(path, "wb") as f:
        for name, a in reads:
            f.write(b">" + name.encode() + b"\n" + a.tobytes() + b"\n")


def write_fastq(path, reads, qual=b"I"):
    with open(path, "wb") as f:
        for name, a in reads:
            f.write(b"@" + name.encode() + b"\n" + a.tobytes() + b"\n+\n" + qual * len(a) + b"\n")


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    g = sub.add_parser("genomes")
    g.add_argument("--out", required=True)
    g.add_argument("--genera", type=int, default=10)
    g.add_argument("--species", type=int, default=10)
    g.add_argument("--len", type=int, default=1000000)
    g.add_argument("--seed", type=int, default=12345)
    g.add_argument("--div", type=float, default=0.03)
    g.add_argument("--cid", action="store_true")
    g.add_argument("--strains", action="store_true")
    r = sub.add_parser("reads")
    r.add_argument("--genomes-dir", required=True)
    r.add_argument("--genera", type=int, default=10)
    r.add_argument("--species", type=int, default=10)
    r.add_argument("--len", type=int, default=1000000)
    r.add_argument("--gseed", type=int, default=12345)
    r.add_argument("--div", type=float, default=0.03)
    r.add_argument("--n", type=int, default=10000)
    r.add_argument("--rdlen", type=int, default=100)
    r.add_argument("--lens", type=int, nargs=2, default=None)
    r.add_argument("--seed", type=int, default=777)
    r.add_argument("--paired", action="store_true")
    r.add_argument("--fastq", action="store_true")
    r.add_argument("--out", required=True, help="output prefix")
    a = ap.parse_args()
    if a.cmd == "genomes":
        write_genomes(a.out, a.genera, a.species, a.len, a.seed, a.div, a.cid, a.strains)
    else:
        seqs = make_genomes(a.genera, a.species, a.len, a.gseed, a.div)
        w = write_fastq if a.fastq else write_fasta
        ext = ".fq" if a.fastq else ".fa"
        if a.paired:
            prs = sample_pairs(seqs, a.n, a.rdlen, a.seed)
            w(a.out + "_1" + ext, [(n, x) for n, x, _ in prs])
            w(a.out + "_2" + ext, [(n, y) for n, _, y in prs])
        else:
            w(a.out + ext, sample_reads(seqs, a.n, a.rdlen, a.seed, lens=a.lens))


if __name__ == "__main__":
    sys.exit(main())
