#!/usr/bin/env python
"""Summarises an .ncu-rep (read on the CPU box with `ncu -i`): one block of key metrics per profiled launch.
usage: tools/ncu_summary.py <report.ncu-rep> [substring ...]   (extra substrings select more metrics)"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
        "gpu__dram_throughput", "lts__t_bytes.sum", "lts__throughput", "l1tex__throughput", "sm__throughput.avg.pct",
        "sm__warps_active.avg.pct", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "launch__occupancy_limit", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "smsp__average_warp", "smsp__warps_eligible",
        "sm__inst_executed_pipe", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg ", "sm__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    extra = sys.argv[2:]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=== launch %s  %s  grid %s block %s" % (r[0], r[hdr.index("Kernel Name")][:70], r[hdr.index("Grid Size")],
                                                       r[hdr.index("Block Size")]))
        for i, h in enumerate(hdr):
            if any(k in h for k in KEYS + extra) or "issue_stalled" in h and "per_warp_active" in h and "not_issued" not in h:
                if r[i] not in ("", "0", "n/a"):
                    print("  %-95s %s %s" % (h, r[i], units[i]))


if __name__ == "__main__":
    main()
