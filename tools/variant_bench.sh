#!/bin/bash
# usage: tools/variant_bench.sh v1 v2 ...   (on the GPU box) -- parity spot check + ncu time/instructions of the tile kernel
for v in "$@"; do
  lib=$PWD/image_transformation_b200/_lib/variants/$v.so
  ok=$(B200COMP_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "batch or golden" 2>&1 | tail -1)
  r=$(B200COMP_LIB=$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:"stream" -s 3 -c 2 --csv python bench.py --batch 16 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | grep -E "gpu__time|inst_exec" | awk -F"\",\"" '{printf "%s ", $NF}' | tr -d '"')
  echo "== $v : $r | $ok"
done
