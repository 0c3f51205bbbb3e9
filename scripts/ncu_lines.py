#!/usr/bin/env python
"""Aggregate an ncu report's source page per CUDA source line: instructions executed and stall samples.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    if r[0] != "":                      # a source line row: aggregated metrics for the line
        try:
            ie = hdr.index("Instructions Executed")
            ss = hdr.index("Warp Stall Sampling (All Samples)")
            key = (cur_file, int(r[0]), r[1].strip()[:90])
            inst = float(r[ie]) if r[ie] not in ("-", "") else 0.0
            samp = float(r[ss]) if r[ss] not in ("-", "") else 0.0
            a = lines.setdefault(key, [0.0, 0.0])
            a[0] += inst
            a[1] += samp
        except (ValueError, IndexError):
            pass
tot_i = sum(v[0] for v in lines.values()) or 1
tot_s = sum(v[1] for v in lines.values()) or 1
print(f"total warp-instructions {tot_i:.0f}, stall samples {tot_s:.0f}")
print("--- by instructions executed")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*v[0]/tot_i:5.1f}% inst {100*v[1]/tot_s:5.1f}% samp  {k[0]}:{k[1]:<4d} {k[2]}")
print("--- by stall samples")
for k, v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*v[1]/tot_s:5.1f}% samp {100*v[0]/tot_i:5.1f}% inst  {k[0]}:{k[1]:<4d} {k[2]}")
