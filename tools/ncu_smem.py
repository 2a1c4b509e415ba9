#!/usr/bin/env python
"""Shared-memory wavefronts per source line: usage tools/ncu_smem.py report.ncu-rep [top]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; lines = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":
        d = dict(zip(hdr[4:], r[4:]))
        g = lambda k: int(d.get(k) or 0)
        lines.append((cur, int(r[0]), r[1].strip()[:80], g("L1 Wavefronts Shared"), g("L1 Wavefronts Shared Excessive"), g("L1 Wavefronts Shared Ideal"), g("Instructions Executed")))
tw = sum(l[3] for l in lines); te = sum(l[4] for l in lines)
print(f"total shared wavefronts {tw}, excessive {te}")
for l in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{l[3]/tw*100:5.1f}% wf ({l[4]/max(tw,1)*100:4.1f}% exc, ideal {l[5]})  inst {l[6]}  {l[0]}:{l[1]}  {l[2]}")
