#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per kernel:
the CUDA source lines with the most stall samples / executed warp instructions.

    ncu -i gpurun_out/prof.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src.csv
    python tools/ncu_hot_lines.py /tmp/src.csv [top_n] > profiles/rXX_ncu_hot_lines.txt
"""
import collections
import csv
import os
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 14
    kernels = collections.OrderedDict()
    func = fname = None
    hdr = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            func = r[1].split("(")[0]; continue
        if r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}; continue
        if hdr is None or r[0] == "":
            continue
        try:
            smp = int(r[hdr["# Samples"]]); ins = int(r[hdr["Instructions Executed"]])
        except (ValueError, KeyError):
            continue
        stalls = {k[6:]: int(r[i]) for k, i in hdr.items() if k.startswith("stall_") and "(" not in k and r[i].isdigit()}
        kernels.setdefault(func, []).append((smp, ins, fname, r[0], r[1].strip(), stalls))
    for func, lines in kernels.items():
        ts = sum(l[0] for l in lines) or 1
        ti = sum(l[1] for l in lines) or 1
        agg = collections.Counter()
        for l in lines:
            agg.update(l[5])
        tot = sum(agg.values()) or 1
        print("== %s: %d samples, %d warp-instructions" % (func, ts, ti))
        print("   stall mix: " + ", ".join("%s %.0f%%" % (k, 100. * v / tot) for k, v in agg.most_common(7)))
        for smp, ins, fname, ln, src, st in sorted(lines, key=lambda l: -max(l[0] / ts, l[1] / ti))[:top]:
            print("   %5.1f%% smp %5.1f%% ins  %s:%s  %s" % (100. * smp / ts, 100. * ins / ti, fname[:14], ln, src[:110]))


if __name__ == "__main__":
    main()
