#!/usr/bin/env python3
"""Map ncu's per-SASS-instruction counters onto source lines.
usage: sass_hotspots.py <ncu source-page csv (sass)> <nvdisasm -g -c output> <kernel name substring> [top]"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass_path, kname = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: k for k, h in enumerate(hdr)}
    prof = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        prof.append((r[ix["Source"]].strip(), float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0),
                     float(r[ix["Warp Stall Sampling (Not-issued Samples)"]] or 0)))
    # nvdisasm: collect (line, opcode text) for the kernel's section, in order
    lines = open(sass_path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l or (l.strip().endswith(":") and kname in l and ".text" in l))
    cur = ("?", 0)
    sass = []
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith("//--------------------- .text"):
            if sass:
                break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            sass.append((cur, m.group(2).strip()))
    print(f"profiled instrs {len(prof)}, disassembled {len(sass)}")
    n = min(len(prof), len(sass))
    by_line = defaultdict(lambda: [0.0, 0.0, 0.0])
    by_op = defaultdict(float)
    tot_i = sum(p[1] for p in prof)
    tot_s = sum(p[2] for p in prof)
    for k in range(n):
        (f, ln), op = sass[k]
        by_line[(f, ln)][0] += prof[k][1]
        by_line[(f, ln)][1] += prof[k][2]
        by_line[(f, ln)][2] += prof[k][3]
        by_op[prof[k][0].split()[0] if not prof[k][0].startswith("@") else prof[k][0].split()[1]] += prof[k][1]
    print(f"total warp-instr {tot_i:.3e}, samples {tot_s:.0f}")
    print("--- by source line (instr %, samples %, stalled samples %)")
    for (f, ln), (i, s, st) in sorted(by_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{f}:{ln:<5d} inst {100 * i / tot_i:5.1f}%  samples {100 * s / tot_s:5.1f}%  not-issued {100 * st / tot_s:5.1f}%")
    print("--- by opcode")
    for op, i in sorted(by_op.items(), key=lambda kv: -kv[1])[:25]:
        print(f"{op:24s} {100 * i / tot_i:5.1f}%")


if __name__ == "__main__":
    main()
