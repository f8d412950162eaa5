#!/usr/bin/env python3
"""Instruction / stall-sample share of each phase of composite_stream_kernel.
usage: region_breakdown.py <ncu sass source csv> <nvdisasm -g -c dump> <tile_kernel.cuh>"""
import csv
import re
import sys
from collections import defaultdict


def main():
    src_csv, sass_path, cuh = sys.argv[1:4]
    rows = list(csv.reader(open(src_csv)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    ix = {h: k for k, h in enumerate(hdr)}
    prof = [(r[ix["Source"]].strip(), float(r[ix["Instructions Executed"]] or 0), float(r[ix["# Samples"]] or 0))
            for r in rows[hi + 1:] if len(r) >= len(hdr)]
    lines = open(sass_path).read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and "composite_stream_kernel" in l)
    cur = ("?", 0)
    sass = []
    for l in lines[start + 1:]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l):
            sass.append(cur)
        if len(sass) >= len(prof):
            break
    src = open(cuh).read().split("\n")

    def find(s):
        return next(i + 1 for i, l in enumerate(src) if s in l)

    kernel = "composite_stream_kernel"
    marks = [(find("void mbar_init"), "mbarrier waits (patch / background landed)"),
             (find("void tma_load_patch"), "TMA / bulk-group issue helpers"),
             (find("void prefetch_l1"), "prefetch"),
             (find("void tile_hpass("), "hpass"), (find("void tile_vpass_over("), "vpass"),
             (find("struct DevPlacementT"), "binning (not this kernel)"),
             (find("composite_stream_kernel(const Cmd"), "main:setup / ring prologue"),
             (find("auto producer_advance"), "main:producer (TMA issue)"),
             (find("auto finish_tile"), "main:finish tile (store)"),
             (find("for (;;) {"), "main:ring + dispatch"),
             (find("if (ring[pos & (kRing - 1)].w[0] == kCmdTile) {"), "main:tile begin"),
             (find("if (kind == kCmdResample) {"), "main:resample glue (decode, waits)"),
             (find("__syncthreads();  // (B) H pass done"), "main:barrier (B) + V pass call"),
             (find("} else if (kind == kCmdIdentTma) {"), "main:identity over"),
             (find("if (--steps_left == 0) finish_tile();"), "main:step end")]

    def region(f, ln):
        if f != cuh.split("/")[-1] or ln < marks[0][0]:
            return None
        r = None
        for start_ln, name in marks:
            if ln >= start_ln - 1:
                r = name
        return r

    agg = defaultdict(lambda: [0.0, 0.0])
    last = "?"
    for (f, ln), (_, ie, sm) in zip(sass, prof):
        r = region(f, ln)
        if r is None:
            r = last
        else:
            last = r
        agg[r][0] += ie
        agg[r][1] += sm
    tot = sum(p[1] for p in prof)
    tots = sum(p[2] for p in prof)
    print(f"total warp-instr {tot / 1e6:.1f}M, samples {tots:.0f}")
    for r, (ie, sm) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{r:24s} inst {100 * ie / tot:5.1f}% ({ie / 1e6:7.1f}M)  samples {100 * sm / tots:5.1f}%")


if __name__ == "__main__":
    main()
