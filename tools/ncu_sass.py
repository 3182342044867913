#!/usr/bin/env python
"""Per-SASS-instruction view of one kernel from `ncu --page source --csv --print-source cuda,sass`:
address order, executed warp instructions, avg active threads, stall samples, CUDA line.

    python tools/ncu_sass.py /tmp/src.csv <kernel substring> [min share of the kernel's instructions, %]
"""
import csv
import os
import sys


def main():
    path, want = sys.argv[1], sys.argv[2]
    floor = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    func = fname = None
    line = -1
    hdr = None
    rows = {}
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = os.path.basename(r[1]); continue
        if r[0] == "Function Name":
            func = r[1]; continue
        if r[0] == "Line No":
            hdr = r; continue
        if hdr is None or want not in (func or ""):
            continue
        if r[0] != "":
            try:
                line = int(r[0])
            except ValueError:
                line = -1
            continue
        ia = hdr.index("Address")
        addr, sass = r[ia], r[ia + 1]
        if not addr:
            continue
        g = {h: r[i] for i, h in enumerate(hdr) if i > ia + 1}
        key = (func, addr)
        def num(x):
            try:
                return float(x.replace(",", ""))
            except ValueError:
                return 0.
        rows[key] = (addr, sass, int(num(g["Instructions Executed"])), num(g["Avg. Threads Executed"]),
                     int(num(g["# Samples"])), "%s:%d" % (fname, line))
    tot = sum(v[2] for v in rows.values()) or 1
    smp = sum(v[4] for v in rows.values()) or 1
    print("# %s: %d SASS instructions, %d executed warp instructions, %d samples" % (want, len(rows), tot, smp))
    for key in sorted(rows, key=lambda k: int(k[1], 16) if k[1].startswith("0x") or all(c in "0123456789abcdef" for c in k[1]) else 0):
        addr, sass, n, thr, s, loc = rows[key]
        if 100. * n / tot < floor and 100. * s / smp < floor:
            continue
        print("%6s %5.2f%% ins %5.2f%% smp thr %4.1f  %-60s %s" % (addr[-5:], 100. * n / tot, 100. * s / smp, thr, sass[:60], loc))


if __name__ == "__main__":
    main()
