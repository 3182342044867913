#!/usr/bin/env python
"""DRAM bytes and executed warp instructions per PSM of every kernel in an `ncu --page raw --csv` export
-> JSON fragment for profiles/traffic.json.
    python tools/ncu_traffic.py raw.csv <psms per launch>"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = float(sys.argv[2])
hdr, units, body = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in body:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    tot = 0.
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ix[m]].replace(",", "")) * scale[units[ix[m]]]
    key = ("bin_topn" if name.startswith("k_bin_") else "select" if name.startswith("k_select") else
           "count_score" if name.startswith("k_count_score") else "ascore" if name.startswith("k_ascore") else name)
    e = out.setdefault(key, {"dram_bytes_per_psm": 0., "warp_inst_per_psm": 0., "kernels": []})
    e["dram_bytes_per_psm"] += tot / n
    e["warp_inst_per_psm"] += float(r[ix["smsp__inst_executed.sum"]].replace(",", "")) / n
    e["kernels"].append(name)
print(json.dumps(out, indent=1))
