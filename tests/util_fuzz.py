"""Seeded mutators that turn a handful of clean reads into FASTQ / FASTA text with the irregularities real files
have (and some they should not): wrapped lines, CR and CR-LF line ends, blank lines, comments, missing final line
end, empty names and reads, lower case, '.', digits, '-', '*', extra '>' in names, '+' lines with text, one spare
quality value.  Test infrastructure."""
import lzma
import os

import util


def clean_reads():
    recs, cur = [], None
    with lzma.open(os.path.join(util.GOLDEN, "adv.reads.fa.xz")) as f:
        for l in f.read().split(b"\n"):
            if l[:1] == b">":
                cur = [l[1:], b""]; recs.append(cur)
            elif cur is not None:
                cur[1] += l
    return [(n, s) for n, s in recs if len(s) >= 30]


def mutate_fastq(rng, recs):
    out = []
    for n, s in recs:
        q = bytes(rng.randrange(33, 74) for _ in s)
        name, seq, plus, qual = b"@" + n, s, b"+", q
        r = rng.random()
        if r < 0.06:
            seq = seq[:10] + b"\n" + seq[10:]
            qual = qual[:10] + b"\n" + qual[10:] if rng.random() < 0.3 else qual      # a wrapped quality line is an error
        elif r < 0.10: name += b" extra\tstuff"
        elif r < 0.14: name += b"/1"
        elif r < 0.18: seq = seq.lower()
        elif r < 0.22: seq = seq[:5] + b"." + seq[6:]
        elif r < 0.25: seq = seq[:5] + b"12" + seq[5:]
        elif r < 0.28: seq = seq[:5] + b"-*" + seq[5:]
        elif r < 0.31: plus = b"+" + n
        elif r < 0.34: name = b"@"
        elif r < 0.37: seq = b""; qual = b""
        elif r < 0.40: qual = qual + b"I"
        elif r < 0.42: seq = seq[:1]; qual = qual[:1]
        elif r < 0.44: seq = b"N" * len(seq)
        rec = name + b"\n" + seq + b"\n" + plus + b"\n" + qual + b"\n"
        r = rng.random()
        if r < 0.05: rec = rec.replace(b"\n", b"\r\n")
        elif r < 0.09: rec = b"\n" + rec
        elif r < 0.12: rec = rec + b"\n\n"
        elif r < 0.14: rec = rec.replace(b"\n", b"\r", 1)
        out.append(rec)
    data = b"".join(out)
    if rng.random() < 0.3: data = data.rstrip(b"\r\n")
    if rng.random() < 0.1: data = b"\n\n" + data
    return data


def mutate_fasta(rng, recs):
    out = []
    for n, s in recs:
        name, seq = b">" + n, s
        r = rng.random()
        if r < 0.08: seq = b"\n".join(seq[i:i + 25] for i in range(0, len(seq), 25))
        elif r < 0.12: name += b" desc words"
        elif r < 0.15: name += b">x"
        elif r < 0.18: seq = seq.lower()
        elif r < 0.22: seq = seq[:5] + b".-*1" + seq[5:]
        elif r < 0.25: name = b">"
        elif r < 0.28: seq = b""
        elif r < 0.30: seq = seq[:1]
        elif r < 0.33: name = b";comment line\n" + name
        elif r < 0.36: name = b"#another\n" + name
        rec = name + b"\n" + seq + b"\n"
        r = rng.random()
        if r < 0.05: rec = rec.replace(b"\n", b"\r\n")
        elif r < 0.09: rec = b"\n" + rec                 # at the head of the file: the reference swallows the header line
        elif r < 0.12: rec = rec + b"\n"
        out.append(rec)
    data = b"".join(out)
    if rng.random() < 0.3: data = data.rstrip(b"\r\n")
    return data
