#!/bin/bash
# usage (on the GPU box): tools/gpu_prof.sh <tag> [batch]  -- quick parity subset, bench, then ncu --set full of the tile kernel
TAG=${1:-x}; B=${2:-256}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/gputest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/gputest_$TAG.log
CMD="python bench.py --batch $B --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err || { echo bench failed; tail -5 gpurun_out/bench_$TAG.err; exit 1; }
python - <<PY
import json; d=json.load(open("gpurun_out/bench_$TAG.json")); print("ms_per_step", d["ms_per_step"], "per canvas us", 1e3*d["ms_per_step"]/d["config"]["canvases_per_gpu_per_step"], "frac", d["roofline"]["frac"], d["roofline"]["kernel_split"])
PY
ncu --set full --clock-control none --import-source on -k regex:composite_slab -s 3 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log
