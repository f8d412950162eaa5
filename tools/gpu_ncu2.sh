#!/bin/bash
# usage (on the GPU box): tools/gpu_ncu2.sh <variant>...  -- ncu --set full of the tile kernel for each build variant (batch 256)
for v in "$@"; do
  export B200COMP_LIB=$PWD/image_transformation_b200/_lib/variants/$v.so
  CMD="python bench.py --batch 256 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
  $CMD > gpurun_out/plain_$v.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:composite_slab -s 3 -c 1 -o gpurun_out/prof_$v $CMD > gpurun_out/ncu_$v.log 2>&1
  tail -1 gpurun_out/ncu_$v.log
done
