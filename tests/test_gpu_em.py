"""Abundance EM on the device: `cfb_em_abundance` must produce the very doubles of the
reference's sequential loops (aln_sink.h:196-495).  The checker here is a plain-Python restatement of those loops
(Python floats are IEEE doubles and the loops add in the same order), plus the reference's golden reports through
the CLI with the device EM forced on."""
import ctypes as C
import lzma
import os
import subprocess

import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu


from util_em import em_python, random_problem


@pytest.mark.parametrize("seed,n,K", [(1, 7, 30), (2, 60, 400), (3, 300, 3000), (4, 5, 3)])
def test_device_em_is_bit_identical_to_the_sequential_loops(seed, n, K):
    count, key_off, target, length, p0 = random_problem(seed, n, K)
    want, want_it, want_diff = em_python(count, key_off, target, length, list(p0))
    lib = C.CDLL(os.path.join(util.ROOT, "centrifuge_b200", "libcfb200.so"))
    a_count = np.array(count, dtype=np.uint64); a_off = np.array(key_off, dtype=np.uint64); a_tgt = np.array(target, dtype=np.uint32)
    a_len = np.array(length, dtype=np.uint64); a_p = np.array(p0, dtype=np.float64)
    iters = C.c_uint64(); diff = C.c_double()
    ptr = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.cfb_em_abundance(C.c_int(0), C.c_uint64(n), C.c_uint64(K), ptr(a_count, C.c_uint64), ptr(a_off, C.c_uint64), ptr(a_tgt, C.c_uint32),
                              ptr(a_len, C.c_uint64), ptr(a_p, C.c_double), C.byref(iters), C.byref(diff))
    assert rc == 0
    assert iters.value == want_it
    assert np.array_equal(a_p.view(np.uint64), np.array(want, dtype=np.float64).view(np.uint64))
    assert diff.value == want_diff


@pytest.mark.parametrize("case,opts", [("default", []), ("k50", ["-k", "50"]), ("family", ["--classification-rank", "family"])])
def test_cli_report_with_device_em_matches_reference_golden(case, opts, tmp_path):
    base = util.golden_index("adv")
    reads = str(tmp_path / "reads.fa")
    with lzma.open(os.path.join(util.GOLDEN, "adv.reads.fa.xz")) as f, open(reads, "wb") as g:
        g.write(f.read())
    exe = os.path.join(util.ROOT, "centrifuge_b200", "centrifuge-class")
    outs = {}
    for mode in ("1", "0"):
        p = subprocess.run([exe, "-f", "-x", base, "-U", reads, "-S", str(tmp_path / "o.tsv"), "--report-file", str(tmp_path / "o.rep")] + opts,
                           stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, env=dict(os.environ, CFB_GPU_EM=mode))
        assert p.returncode == 0
        with open(tmp_path / "o.rep", "rb") as f:
            rep = f.read()
        outs[mode] = (rep, [l for l in p.stderr.decode().splitlines() if "EM algorithm" in l or "Probability diff" in l])
        with open(os.path.join(util.GOLDEN, "adv.%s.report.tsv" % case), "rb") as f:
            assert rep == f.read()
    assert outs["1"] == outs["0"]          # same iteration count and final difference as the host loop
