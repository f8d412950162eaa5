#!/bin/bash
# usage (on the GPU box): tools/gpu_evidence.sh  -- the ncu evidence of the round: launch lists (C3 batch 256, C5 batch 16) and one
# --set full capture of the tile kernel at batch 256; every command first runs without ncu.
set -x
C3="python bench.py --batch 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan"
C5="python bench.py --workload c5_8k_64obj --batch 16 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan"
$C3 > gpurun_out/ev_c3_bench.json 2> gpurun_out/ev_c3.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches.csv $C3 > /dev/null 2>&1
$C5 > gpurun_out/ev_c5_bench.json 2> gpurun_out/ev_c5.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/r2_launches_c5.csv $C5 > /dev/null 2>&1
$C3 > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite_slab -s 3 -c 1 -o gpurun_out/prof_r2_final $C3 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
