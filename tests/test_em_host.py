"""Abundance EM, host form (`cfb_em_abundance_host`, the loop `centrifuge-class` runs for small tie-set tables): the very
doubles and iteration count of the reference's sequential loops, checked against a plain-Python restatement."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from util_em import em_python, random_problem


@pytest.mark.parametrize("seed,n,K", [(1, 7, 30), (2, 60, 400), (3, 300, 3000), (4, 5, 3), (5, 40, 1)])
def test_host_em_is_bit_identical_to_the_sequential_loops(seed, n, K):
    count, key_off, target, length, p0 = random_problem(seed, n, K)
    want, want_it, want_diff = em_python(count, key_off, target, length, list(p0))
    lib = C.CDLL(util.PRODUCT_LIB)
    a_count = np.array(count, dtype=np.uint64); a_off = np.array(key_off, dtype=np.uint64); a_tgt = np.array(target, dtype=np.uint32)
    a_len = np.array(length, dtype=np.uint64); a_p = np.array(p0, dtype=np.float64)
    iters = C.c_uint64(); diff = C.c_double()
    ptr = lambda a, t: a.ctypes.data_as(C.POINTER(t))
    rc = lib.cfb_em_abundance_host(C.c_uint64(n), C.c_uint64(K), ptr(a_count, C.c_uint64), ptr(a_off, C.c_uint64), ptr(a_tgt, C.c_uint32),
                                   ptr(a_len, C.c_uint64), ptr(a_p, C.c_double), C.byref(iters), C.byref(diff))
    assert rc == 0
    assert iters.value == want_it
    assert np.array_equal(a_p.view(np.uint64), np.array(want, dtype=np.float64).view(np.uint64))
    assert diff.value == want_diff
