#!/bin/bash
# usage (GPU box): tools/ab_split.sh default variant ...  -- per-phase ms of one plan run (batch 256) per build variant
for v in "$@"; do
  if [ "$v" = "default" ]; then unset B200COMP_LIB; else export B200COMP_LIB=$PWD/image_transformation_b200/_lib/variants/$v.so; fi
  r=$(timeout 120 python bench.py --batch 256 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_split']; print('step %.3f ms  tile %.3f  binning %.3f  prepare %.3f' % (d['ms_per_step'], k['tile_kernel_ms_per_step'], k['binning_ms_per_step'], k['prepare_ms_per_step']))")
  echo "== $v : $r"
done
