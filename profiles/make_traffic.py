#!/usr/bin/env python3
"""profiles/ncu_traffic.json from an ncu launch list of `bench.py --batch N --steps 2 --warmup 3 --no-e2e
--no-cpu-baseline` (metrics gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum).

usage: make_traffic.py <launches.csv> <bench json of the same command> <out json>

The LAST step of the run is used: the launches from the last prepare_cutouts_kernel to the last
composite_slab_kernel (prepare + 3 binning kernels + tile kernel = one b200comp_plan_run)."""
import csv
import json
import sys


def main():
    src, bench_json, out = sys.argv[1:4]
    rows = [r for r in csv.reader(open(src)) if len(r) > 14 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"name": r[4].split("(")[0]})
        d[r[12]] = float(r[14].replace(",", ""))
    ids = sorted(launches)
    last_tile = max(i for i in ids if launches[i]["name"].startswith("composite_slab_kernel"))
    first = max(i for i in ids if i < last_tile and launches[i]["name"].startswith("prepare_cutouts_kernel"))
    step = [launches[i] for i in ids if first <= i <= last_tile]
    dram = sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in step)
    t = sum(l["gpu__time_duration.sum"] for l in step)
    # a run over many canvases is cut into waves: several launches of every kernel but the prepare pass per step
    tiles = [l for l in step if l["name"].startswith("composite_slab_kernel")]
    by_name = {}
    for l in step:
        by_name[l["name"]] = by_name.get(l["name"], 0.0) + l["gpu__time_duration.sum"]
    names = []
    for l in step:
        if l["name"] not in names:
            names.append(l["name"])
    bench = json.loads(open(bench_json).read().strip().splitlines()[-1])
    res = {
        "source": f"profiles/{src.split('/')[-1]} (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                  f"--clock-control none; last plan run of bench.py --batch {bench['config']['canvases_per_gpu_per_step']} "
                  f"-- the phase-split pass, wave split off, kernels one after another: "
                  + " + ".join(f"{sum(1 for l in step if l['name'] == n)} x {n}" for n in names) + ")",
        "dram_bytes_per_launch": dram,
        "algorithmic_bytes_per_launch": bench["roofline"]["algorithmic_bytes_per_launch"],
        "tile_kernel_dram_bytes": sum(l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"] for l in tiles),
        "tile_kernel_share_of_step_time": sum(l["gpu__time_duration.sum"] for l in tiles) / t,
        "launch_ns": by_name,
        "launches_per_step": len(step),
    }
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
