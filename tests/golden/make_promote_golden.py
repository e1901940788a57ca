#!/usr/bin/env python3
"""Golden outputs of the reference's own `centrifuge-promote` (Perl) on the committed golden classification outputs
and on a crafted table of corner cases.  The script and a one-line `centrifuge-inspect` shim (-> oracle/_ref/
centrifuge-inspect-bin) are staged in a temporary directory because the script looks for centrifuge-inspect next to
itself; nothing from /root/reference enters the repository.  Run in the build container (needs perl and /root/reference)."""
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

LEVELS = ["genus", "species", "family", "lca", "phylum"]
CASES = ["default", "k50", "host", "family"]


def stage():
    tmp = tempfile.mkdtemp()
    shutil.copy("/root/reference/centrifuge-promote", os.path.join(tmp, "centrifuge-promote"))
    shim = os.path.join(tmp, "centrifuge-inspect")
    with open(shim, "w") as f:
        f.write("#!/bin/sh\nexec %s \"$@\"\n" % os.path.join(ROOT, "oracle", "_ref", "centrifuge-inspect-bin"))
    os.chmod(shim, os.stat(shim).st_mode | stat.S_IEXEC)
    return tmp


def quirk_rows():
    # equal consecutive readIDs, a re-used name later on, dotted strain taxIDs, taxIDs outside the tree, unclassified rows,
    # rows whose taxID is already at / above the requested level, the root, duplicate promotions within one read
    return ["readID\tseqID\ttaxID\tscore\t2ndBestScore\thitLength\tqueryLength\tnumMatches",
            "a\tcid5\t1005\t400\t0\t40\t100\t2", "a\tcid9\t1009\t400\t0\t40\t100\t2",
            "b\tcid5\t1005\t900\t0\t60\t100\t3", "b\tcid6\t1006\t900\t0\t60\t100\t3", "b\tcid10\t1010\t900\t0\t60\t100\t3",
            "c\tcid5\t1005.7\t900\t0\t60\t100\t1", "d\tx\t99999\t900\t0\t60\t100\t1",
            "e\tunclassified\t0\t0\t0\t0\t100\t1", "e\tunclassified\t0\t0\t0\t0\t100\t1",
            "f\tgenus\t100\t100\t0\t20\t100\t2", "f\tcid2\t1002\t2500\t0\t65\t100\t2",
            "g\tfamily\t10\t900\t0\t60\t100\t1", "h\tno rank\t1\t900\t0\t60\t100\t1",
            "a\tcid3\t1003\t900\t0\t60\t100\t1",
            "i\tgenus\t103\t900\t0\t60\t100\t2", "i\tfamily\t11\t900\t0\t60\t100\t2"]


def main():
    tmp = stage()
    base = util.golden_index("adv")
    for case in CASES:
        tsv = os.path.join(tmp, case + ".tsv")
        with lzma.open(os.path.join(HERE, "adv.%s.tsv.xz" % case)) as f, open(tsv, "wb") as g:
            g.write(f.read())
        for lv in (LEVELS if case == "default" else ["genus", "lca"]):
            out = subprocess.run(["perl", os.path.join(tmp, "centrifuge-promote"), base, tsv, lv], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
            with lzma.open(os.path.join(HERE, "adv.%s.promote.%s.tsv.xz" % (case, lv)), "wb", preset=9) as g:
                g.write(out)
    tsv = os.path.join(HERE, "promote_quirks.tsv")
    with open(tsv, "w") as f:
        f.write("\n".join(quirk_rows()) + "\n")
    for lv in LEVELS:
        out = subprocess.run(["perl", os.path.join(tmp, "centrifuge-promote"), base, tsv, lv], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, check=True).stdout
        with open(os.path.join(HERE, "promote_quirks.%s.tsv" % lv), "wb") as g:
            g.write(out)
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
