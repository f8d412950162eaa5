#!/bin/bash
# usage (on the GPU box): tools/gpu_evidence.sh  -- the ncu evidence of the round: launch lists (C3 batch 256, C4 batch 64, C5 batch 16,
# aux kernels) and one --set full capture of the tile kernel at batch 256 with the wave split off (one 256-canvas launch); every
# command first runs without ncu.
set -x
mkdir -p gpurun_out
LL="ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv"
C3="python bench.py --batch 256 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan"
C4="python bench.py --workload c4_aspect_sweep --batch 64 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan"
C5="python bench.py --workload c5_8k_64obj --batch 16 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan"
AUX="python profiles/aux_kernels_bench.py"
$C3 > gpurun_out/ev_c3_bench.json 2> gpurun_out/ev_c3.err && $LL --log-file gpurun_out/r2_launches.csv $C3 > /dev/null 2>&1
$C4 > gpurun_out/ev_c4_bench.json 2> gpurun_out/ev_c4.err && $LL --log-file gpurun_out/r2_launches_c4.csv $C4 > /dev/null 2>&1
$C5 > gpurun_out/ev_c5_bench.json 2> gpurun_out/ev_c5.err && $LL --log-file gpurun_out/r2_launches_c5.csv $C5 > /dev/null 2>&1
$AUX > gpurun_out/r2_aux_kernels.json 2> gpurun_out/ev_aux.err && $LL --log-file gpurun_out/r2_aux_launches.csv $AUX > /dev/null 2>&1
if [ -z "$SKIP_FULL" ]; then
  B200COMP_WAVES=1 $C3 > /dev/null 2>&1 && \
  B200COMP_WAVES=1 ncu --set full --clock-control none --import-source on -k regex:composite_slab -s 3 -c 1 -o gpurun_out/prof_r2_final $C3 > gpurun_out/ncu_final.log 2>&1
  tail -2 gpurun_out/ncu_final.log
fi
ls -la gpurun_out/r2_*
