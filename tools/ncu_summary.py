#!/usr/bin/env python
"""Per-kernel table of the metrics the roofline discussion uses, from
`ncu -i X.ncu-rep --page raw --csv > raw.csv`.   python tools/ncu_summary.py raw.csv"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    names = [r[idx["Kernel Name"]].split("(")[0].replace("void ", "")[:22] for r in body]
    print("%-74s %-9s %s" % ("metric", "unit", " | ".join("%-22s" % n for n in names)))
    for w in WANT:
        if w not in idx:
            continue
        print("%-74s %-9s %s" % (w[:74], units[idx[w]][:9], " | ".join("%-22s" % r[idx[w]][:22] for r in body)))


if __name__ == "__main__":
    main()
