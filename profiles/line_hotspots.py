#!/usr/bin/env python3
"""Per-source-line executed warp instructions / stall samples from `ncu --page source --print-source cuda,sass --csv`.
usage: line_hotspots.py src_cuda.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = []
fpath = "?"
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
        hdr = None
        continue
    if r[0] == "Line No":
        hdr = r
        # two "Source" columns: the first is the CUDA line, the second the SASS (empty on line rows)
        iL, iS = 0, 1
        iA = r.index("Address")
        iN = r.index("Instructions Executed")
        iP = r.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[iL] != "" and r[iA] in ("", "-"):  # a CUDA source line row (aggregated)
        try:
            out.append((fpath, int(r[iL]), r[iS].strip(), float(r[iN] or 0), float(r[iP] or 0)))
        except ValueError:
            pass
ti = sum(o[3] for o in out)
ts = sum(o[4] for o in out)
print(f"lines {len(out)}, warp instructions {ti:.0f}, samples {ts:.0f}")
for f, ln, src, ni, sp in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{100 * ni / ti:5.2f}% instr {100 * sp / ts:5.2f}% samples  {f}:{ln:<5d} {src[:100]}")
