#!/usr/bin/env python3
"""Rank CUDA source lines of one kernel by warp-stall samples.
usage: ncu_hot_lines.py report.ncu-rep kernel-regex [top]   (needs -lineinfo and --import-source on at capture time)"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:" + kern],
                         capture_output=True, text=True).stdout
    path, hdr, lines, total = None, None, [], 0
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            path = r[1].split("/")[-1]; continue
        if len(r) >= 2 and r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0]:          # a CUDA line (SASS rows have an empty line number)
            i_s = hdr.index("# Samples"); i_ex = hdr.index("Instructions Executed"); i_th = hdr.index("Avg. Threads Executed")
            s = int(r[i_s] or 0); total += s
            lines.append((s, path, r[0], r[i_ex], r[i_th], r[1].strip()))
    lines.sort(key=lambda x: -x[0])
    print("# %s: %d stall samples on %d source lines" % (kern, total, len(lines)))
    for s, p, ln, ex, th, src in lines[:top]:
        print("%5.1f%% %-12s %5s  inst %9s thr/inst %4s  %s" % (100.0 * s / max(total, 1), p, ln, ex, th, src[:130]))


if __name__ == "__main__":
    main()
