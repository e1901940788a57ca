#!/usr/bin/env python3
"""profiles/ artefacts from an ncu report: a readable metric summary per kernel and, optionally, the traffic JSON that
bench.py reads for roofline.traffic.
usage: ncu_summary.py <report.ncu-rep> <out.txt> <title> [--traffic <kernel substring> <units in that launch> <workload_key> <out.json>]"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out_txt, title = sys.argv[1:4]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, check=True).stdout.decode()
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ki = hdr.index("Kernel Name")
    lines = ["# %s" % title, "# ncu --set full --clock-control none --import-source on; one launch per kernel; cold caches, serialised", ""]
    for row in data:
        lines.append("## %s" % row[ki])
        for m in KEEP + sorted(h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            if m in hdr:
                lines.append("%-90s %16s %s" % (m, row[hdr.index(m)], units[hdr.index(m)]))
        lines.append("")
    with open(out_txt, "w") as f:
        f.write("\n".join(lines) + "\n")
    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        kname, nunits, key, out_json = sys.argv[k + 1], int(sys.argv[k + 2]), sys.argv[k + 3], sys.argv[k + 4]
        row = [d for d in data if kname in d[ki]][-1]

        def to_bytes(m):
            v, u = float(row[hdr.index(m)].replace(",", "")), units[hdr.index(m)]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
        with open(out_json, "w") as f:
            json.dump({"kernel": row[ki], "units_in_capture": nunits, "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_unit": (rd + wr) / nunits,
                       "workload_key": key, "file": out_txt, "source": "ncu --set full (%s)" % title}, f, indent=1)
    print("\n".join(lines[:14]))


if __name__ == "__main__":
    main()
