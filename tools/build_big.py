#!/usr/bin/env python3
"""Time the GPU index builder on a synthetic genome set:  build_big.py GENERA SPECIES LEN [outdir]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from centrifuge_b200 import capi
g, s, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
d = sys.argv[4] if len(sys.argv) > 4 else "/tmp/cfb200_big_%d_%d_%d" % (g, s, L)
tax = capi.write_synth_taxonomy(d, g, s, L)
o = capi.build_opts(d + "/idx", synth=(g, s, L, 12345, 0.03), conversion_table=tax[0], taxonomy_tree=tax[1], name_table=tax[2], verbose=1)
t0 = time.time(); capi.build_index(o); t1 = time.time()
sz = sum(os.path.getsize("%s/idx.%d.cf" % (d, k)) for k in (1, 2, 3, 4))
print("built %.3f Gbp in %.1f s -> %.2f GB of .cf files" % (g * s * L / 1e9, t1 - t0, sz / 1e9))
