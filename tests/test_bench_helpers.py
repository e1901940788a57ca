"""bench.py's host-side input builders (no GPU): the byte form, the packed form and the FASTQ text of the same reads
must describe the same reads -- the timed arms are only comparable if they do."""
import os
import sys

import numpy as np

import util

sys.path.insert(0, util.ROOT)
import bench  # noqa: E402


def random_reads(n, lo, hi, mates, seed):
    rng = np.random.default_rng(seed)
    codes = rng.integers(0, 4, size=(mates, n, hi), dtype=np.uint8)
    codes[rng.random((mates, n, hi)) < 0.03] = 4
    lens = rng.integers(lo, hi + 1, size=(mates, n)).astype(np.uint32)
    codes[np.arange(hi)[None, None, :] >= lens[:, :, None]] = 4
    return bench.Reads(codes, lens)


def plain(shape, dt):
    return np.zeros(shape, dtype=dt)


def test_byte_packed_and_text_forms_agree():
    from centrifuge_b200 import capi
    for mates, lo, hi in ((1, 100, 100), (1, 75, 300), (2, 150, 150), (2, 1, 70)):
        rd = random_reads(257, lo, hi, mates, seed=hi + mates)
        bases, offs, lens, fl = rd.byte_form(plain)
        for m in range(mates):
            for i in (0, 1, 100, 256):
                L = int(rd.lens[m][i])
                assert np.array_equal(bases[int(offs[m][i]):int(offs[m][i]) + L], rd.codes[m][i][:L])
                assert lens[m][i] == L
        cb = capi.make_batch(bases, offs[0], lens[0], offs[1] if mates == 2 else None, lens[1] if mates == 2 else None, fl)
        w_ref, n_ref = capi.pack_batch(cb)
        pw, pn, pl, pf = rd.packed_form(plain)
        assert np.array_equal(pw, w_ref) and np.array_equal(np.sort(pn), np.sort(n_ref))
        assert np.array_equal(pf, fl)
        for m in range(mates):
            txt = rd.fastq(m, start=5, suffix=b"/%d" % (m + 1) if mates == 2 else b"").tobytes()
            want = b"".join(b"@r%09d%s\n" % (5 + i, (b"/%d" % (m + 1)) if mates == 2 else b"") + np.frombuffer(b"ACGTN", dtype=np.uint8)[rd.codes[m][i][:rd.lens[m][i]]].tobytes()
                            + b"\n+\n" + b"I" * int(rd.lens[m][i]) + b"\n" for i in range(rd.n))
            assert txt == want
        # the N filter flag: >= 2 bases and at most floor(0.15 len) Ns
        for i in (0, 7, 200):
            for m in range(mates):
                L = int(rd.lens[m][i]); ns = int((rd.codes[m][i][:L] == 4).sum())
                assert ((int(fl[i]) >> m) & 1) == int(L >= 2 and ns <= int(0.15 * L))


def test_argument_defaults_name_the_baseline_configs():
    old = sys.argv
    try:
        sys.argv = ["bench.py"]
        a = bench.parse_args()
        assert (a.genera, a.species, a.genome_len, a.reads, a.lens, a.paired) == (900, 10, 1000000, 10000000, (100, 100), False)
        assert bench.metric_name(a) == "reads/sec (100 bp SE classification)"
        sys.argv = ["bench.py", "--paired", "--rdlen", "150", "--index-gbp", "17"]
        a = bench.parse_args()
        assert a.genera == 1700 and a.reads == 5000000 and a.lens == (150, 150) and bench.metric_name(a) == "pairs/sec (150 bp PE classification)"
        sys.argv = ["bench.py", "--lens", "75-300"]
        a = bench.parse_args()
        assert a.lens == (75, 300) and a.reads == 5000000
    finally:
        sys.argv = old
