#!/usr/bin/env python3
"""Rank the SASS instructions of an ncu source page (--page source --csv) by warp-stall samples, with the dominant
stall reason of each, plus the opcode mix and where long-scoreboard samples concentrate.
usage: stall_hotspots.py src.csv [top]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: k for k, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs = []
for n, r in enumerate(rows[hi + 1:]):
    if len(r) < len(hdr):
        continue
    st = {h: float(r[ix[h]] or 0) for h in stall_cols}
    recs.append((n, r[ix["Source"]].strip(), float(r[ix["# Samples"]] or 0), float(r[ix["Instructions Executed"]] or 0), st))
tot_s = sum(r[2] for r in recs)
tot_i = sum(r[3] for r in recs)
print(f"total samples {tot_s:.0f}, warp instructions {tot_i:.0f}")
agg = Counter()
for _, _, _, _, st in recs:
    for k, v in st.items():
        agg[k] += v
print("stall reasons (share of samples):", ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in agg.most_common(10)))
ops = Counter()
for _, src, _, ni, _ in recs:
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    ops[op.split(".")[0]] += ni
print("opcode mix:", ", ".join(f"{k} {100 * v / tot_i:.1f}%" for k, v in ops.most_common(22)))
print(f"--- top {top} instructions by samples (index, samples %, executed %, dominant stall, sass)")
for n, src, s, ni, st in sorted(recs, key=lambda r: -r[2])[:top]:
    dom = max(st.items(), key=lambda kv: kv[1])
    print(f"{n:6d} {100 * s / tot_s:5.2f}% {100 * ni / tot_i:5.2f}%  {dom[0][6:]:14s} {src[:90]}")
