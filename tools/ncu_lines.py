#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo + --import-source on):
usage: tools/ncu_lines.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; lines = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name": continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":  # a source-line aggregate row
        d = dict(zip(hdr[4:], r[4:]))
        lines.append((cur, int(r[0]), r[1].strip()[:90], int(d["# Samples"] or 0), int(d["Instructions Executed"] or 0)))
ts = sum(l[3] for l in lines); ti = sum(l[4] for l in lines)
print(f"total samples {ts}, instructions {ti}")
print("--- by stall samples")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{l[3]/ts*100:5.1f}% smp {l[4]/ti*100:5.1f}% inst  {l[0]}:{l[1]}  {l[2]}")
print("--- by instructions")
for l in sorted(lines, key=lambda l: -l[4])[:top]:
    print(f"{l[3]/ts*100:5.1f}% smp {l[4]/ti*100:5.1f}% inst  {l[0]}:{l[1]}  {l[2]}")
if len(sys.argv) > 3:
    print("--- ranges")
    for spec in sys.argv[3:]:
        f, a, b = spec.split(":")
        sm = sum(l[3] for l in lines if l[0] == f and int(a) <= l[1] <= int(b))
        ins = sum(l[4] for l in lines if l[0] == f and int(a) <= l[1] <= int(b))
        print(f"{spec}: {sm/ts*100:.1f}% smp {ins/ti*100:.1f}% inst")
