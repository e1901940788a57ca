"""The ctypes binding (centrifuge_b200/capi.py) mirrors the C structs of include/cfb200.h by hand: compile a C probe
against the header and check sizes and field offsets, so the two cannot drift apart silently.  The header must also
compile as plain C (it is the boundary a C, cgo or JNI caller would include)."""
import ctypes as C
import os
import subprocess

import util

PROBE = r"""
#include <stddef.h>
#include <stdio.h>
#include "cfb200.h"
#define S(t) printf(#t " size %zu\n", sizeof(t))
#define F(t, f) printf(#t "." #f " %zu\n", offsetof(t, f))
int main(void) {
  S(cfb_index_info); F(cfb_index_info, line_rate); F(cfb_index_info, device); F(cfb_index_info, device_bytes);
  S(cfb_params); F(cfb_params, host_taxids); F(cfb_params, n_excluded_taxids);
  S(cfb_batch); F(cfb_batch, bases); F(cfb_batch, off); F(cfb_batch, len); F(cfb_batch, flags);
  S(cfb_rec); F(cfb_rec, score); F(cfb_rec, uid);
  S(cfb_result); F(cfb_result, rec_off); F(cfb_result, recs);
  S(cfb_text_opts); F(cfb_text_opts, maxlen_hint);
  S(cfb_text_result); F(cfb_text_result, tsv); F(cfb_text_result, multi); F(cfb_text_result, multi_stride);
  S(cfb_build_opts); F(cfb_build_opts, synth_len); F(cfb_build_opts, synth_div); F(cfb_build_opts, ftab_chars); F(cfb_build_opts, verbose); F(cfb_build_opts, synth_prefix);
  S(cfb_index_tables); F(cfb_index_tables, walk8_bytes); F(cfb_index_tables, walk8_rows); F(cfb_index_tables, ftabd_bytes); F(cfb_index_tables, ftabk_chars); F(cfb_index_tables, ftabd_chars);
  S(cfb_batch_packed); F(cfb_batch_packed, words); F(cfb_batch_packed, len); F(cfb_batch_packed, n_pos); F(cfb_batch_packed, flags);
  S(cfb_synth_read_opts); F(cfb_synth_read_opts, paired); F(cfb_synth_read_opts, ins_hi);
  return 0;
}
"""


def test_ctypes_structs_match_the_c_header(tmp_path):
    from centrifuge_b200 import capi
    src = tmp_path / "probe.c"
    src.write_text(PROBE)
    exe = str(tmp_path / "probe")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(util.ROOT, "include"), "-o", exe, str(src)])
    got = {}
    for line in subprocess.check_output([exe]).decode().splitlines():
        k, *rest = line.split()
        got[k if rest[0] != "size" else k + ".size"] = int(rest[-1])
    import numpy as np
    rec = np.dtype([("taxid", "<u8"), ("score", "<u4"), ("hitlen", "<u4"), ("uid", "<u4"), ("pad", "<u4")])
    pairs = {"cfb_index_info": capi.IndexInfo, "cfb_params": capi.Params, "cfb_batch": capi.BatchC, "cfb_result": capi.ResultC,
             "cfb_text_opts": capi.TextOpts, "cfb_text_result": capi.TextResultC, "cfb_build_opts": capi.BuildOpts,
             "cfb_index_tables": capi.IndexTables, "cfb_batch_packed": capi.BatchPackedC, "cfb_synth_read_opts": capi.SynthReadOpts}
    for cname, st in pairs.items():
        assert got[cname + ".size"] == C.sizeof(st), cname
        for key, off in got.items():
            if key.startswith(cname + ".") and not key.endswith(".size"):
                assert getattr(st, key.split(".", 1)[1]).offset == off, key
    assert got["cfb_rec.size"] == rec.itemsize and got["cfb_rec.score"] == rec.fields["score"][1] and got["cfb_rec.uid"] == rec.fields["uid"][1]


def test_host_packer_matches_a_numpy_restatement():
    """cfb_pack_batch (host code, no device needed) == a vectorised numpy packing of the same reads."""
    import numpy as np
    from centrifuge_b200 import capi
    rng = np.random.default_rng(3)
    n, L = 500, 77
    codes = rng.integers(0, 4, size=(n, L), dtype=np.uint8)
    codes[rng.random((n, L)) < 0.02] = 4
    offs = np.arange(n, dtype=np.uint64) * np.uint64(L); lens = np.full(n, L, dtype=np.uint32)
    b = capi.make_batch(np.ascontiguousarray(codes.reshape(-1)), offs, lens)
    words, npos = capi.pack_batch(b)
    w2, p2 = capi.pack_fixed(codes)
    assert np.array_equal(words, w2) and np.array_equal(npos, p2)
    # unpack in numpy and compare with the input
    W = (L + 31) // 32
    un = ((words.reshape(n, W, 1) >> (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, None, :]) & np.uint64(3)).reshape(n, W * 32)[:, :L].astype(np.uint8)
    un.reshape(-1)[0:0] = 0
    flat = un.copy()
    r = (npos >> np.uint64(5)) // np.uint64(W); j = ((npos >> np.uint64(5)) % np.uint64(W)) * np.uint64(32) + (npos & np.uint64(31))
    flat[r.astype(np.int64), j.astype(np.int64)] = 4
    assert np.array_equal(flat, codes)
