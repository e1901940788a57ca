"""Commit the reference's only known-answer fixture (example/index/test.*.cf + example/reads/input.fa,
MANUAL.markdown:1586-1603) as small compressed files so that GPU boxes -- which have no /root/reference --
can run BASELINE.json configs[0].  Data fixtures only; run in a container that holds the reference tree:

    python tests/golden/make_example_golden.py
"""
import lzma
import os
import shutil

REF = "/root/reference/example"
HERE = os.path.dirname(os.path.abspath(__file__))

for k in "1234":
    with open("%s/index/test.%s.cf" % (REF, k), "rb") as f, lzma.open(os.path.join(HERE, "example.%s.cf.xz" % k), "wb", preset=9) as g:
        g.write(f.read())
shutil.copyfile(REF + "/reads/input.fa", os.path.join(HERE, "example.reads.fa"))
print("wrote example.{1,2,3,4}.cf.xz and example.reads.fa")
