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
    shutil.rmtree(tmp)


if __name__ == "__main__":
    main()
