#!/usr/bin/env python3
"""Golden Kraken-style reports: the reference's own `centrifuge-kreport` (Perl) run on the committed golden
classification outputs.  The script and a one-line `centrifuge-inspect` shim (-> oracle/_ref/centrifuge-inspect-bin)
are staged in a temporary directory because the script looks for centrifuge-inspect next to itself; nothing from
/root/reference enters the repository.  Run in the build container (needs perl and /root/reference)."""
import lzma
import os
import shutil
import stat
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402

CASES = ["default", "k1", "k50", "minhit15", "host", "excl", "family", "notraverse"]
VARIANTS = {"": [], ".zeros": ["--show-zeros"], ".minscore": ["--min-score", "300"], ".minlen": ["--min-length", "40"]}


def main():
    tmp = tempfile.mkdtemp()
    shutil.copy("/root/reference/centrifuge-kreport", os.path.join(tmp, "centrifuge-kreport"))
    shim = os.path.join(tmp, "centrifuge-inspect")
    with open(shim, "w") as f:
        f.write("#!/bin/sh\nexec %s \"$@\"\n" % os.path.join(ROOT, "oracle", "_ref", "centrifuge-inspect-bin"))
    os.chmod(shim, os.stat(shim).st_mode | stat.S_IEXEC)
    base = util.golden_index("adv")
    for case in CASES:
        tsv = os.path.join(tmp, case + ".tsv")
        with lzma.open(os.path.join(HERE, "adv.%s.tsv.xz" % case)) as f, open(tsv, "wb") as g:
            g.write(f.read())
        for suffix, opts in (VARIANTS.items() if case == "default" else [("", [])]):
            out = subprocess.run(["perl", os.path.join(tmp, "centrifuge-kreport"), "-x", base] + opts + [tsv], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            with open(os.path.join(HERE, "adv.%s%s.kreport.txt" % (case, suffix)), "wb") as g:
                g.write(out)
    # crafted rows for the script's corner cases: rows of one read merged to their LCA (also across what the
    # classifier would call two reads with the same name), dotted strain taxIDs and taxIDs outside the tree
    # (both go to the root), unclassified rows, a filtered row between two rows of one read
    rows = ["readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches",
            "a\ts\t1005\t400\t0\t40\t100\t2", "a\ts\t1009\t400\t0\t40\t100\t2",          # same genus -> LCA genus
            "b\ts\t1005\t900\t0\t60\t100\t3", "b\ts\t1006\t900\t0\t60\t100\t3", "b\ts\t1010\t900\t0\t60\t100\t3",
            "c\ts\t1005.7\t900\t0\t60\t100\t1", "d\ts\t99999\t900\t0\t60\t100\t1",
            "e\tunclassified\t0\t0\t0\t0\t100\t1", "e\tunclassified\t0\t0\t0\t0\t100\t1",    # two reads named e
            "f\ts\t1001\t100\t0\t20\t100\t2", "f\ts\t1002\t2500\t0\t65\t100\t2",
            "g\ts\t0\t0\t0\t0\t100\t1", "g\ts\t1003\t900\t0\t60\t100\t1",
            "h\ts\t1\t900\t0\t60\t100\t1", "i\ts\t103\t900\t0\t60\t100\t1", "i\ts\t11\t900\t0\t60\t100\t1"]
    tsv = os.path.join(tmp, "quirks.tsv")
    with open(tsv, "w") as f:
        f.write("\n".join(rows) + "\n")
    shutil.copy(tsv, os.path.join(HERE, "kreport_quirks.tsv"))
    for suffix, opts in VARIANTS.items():
        out = subprocess.run(["perl", os.path.join(tmp, "centrifuge-kreport"), "-x", base] + opts + [tsv], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        with open(os.path.join(HERE, "kreport_quirks%s.kreport.txt" % suffix), "wb") as g:
            g.write(out)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
