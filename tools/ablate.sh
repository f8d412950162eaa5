#!/bin/bash
# usage (GPU box): tools/ablate.sh v1 v2 ...  -- ms/step of measurement builds (no parity tests: their pixels are wrong), waves on and off
for v in "$@"; do
  if [ "$v" = "default" ]; then unset B200COMP_LIB; else export B200COMP_LIB=$PWD/image_transformation_b200/_lib/variants/$v.so; fi
  for w in 1 ""; do
    ms=$(B200COMP_WAVES=$w timeout 120 python bench.py --batch 256 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],2), d['roofline']['kernel_split']['tile_kernel_ms_per_step'])")
    echo "== $v waves='$w': $ms"
  done
done
