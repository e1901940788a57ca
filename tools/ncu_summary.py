#!/usr/bin/env python3
"""profiles/ artefacts from an ncu report: a readable metric summary of one kernel and the traffic JSON bench.py reads.
usage: ncu_summary.py <report.ncu-rep> <kernel substring> <reads in that launch> <out.txt> <out.json> [title]"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, kname, reads, out_txt, out_json = sys.argv[1:6]
    title = sys.argv[6] if len(sys.argv) > 6 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    row = [d for d in data if kname in d[ki]][-1]
    val = lambda m: row[hdr.index(m)]
    lines = ["# %s" % title, "## %s  (ncu --set full --clock-control none --import-source on)" % row[ki].split("(")[0], ""]
    for m in KEEP + sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")):
        if m in hdr:
            lines.append("%-90s %14s %s" % (m, val(m), units[hdr.index(m)]))
    with open(out_txt, "w") as f:
        f.write("\n".join(lines) + "\n")

    def to_bytes(m):
        v, u = float(val(m).replace(",", "")), units[hdr.index(m)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    with open(out_json, "w") as f:
        json.dump({"kernel": kname, "reads_in_capture": int(reads), "dram_bytes_read": rd, "dram_bytes_write": wr,
                   "dram_bytes_per_read": (rd + wr) / int(reads), "source": "%s (ncu --set full, %s)" % (out_txt, title)}, f, indent=1)
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    main()
