#!/bin/bash
# usage (on the GPU box): [WORKLOAD=c3_4k_20obj] [BATCH=128] tools/ab.sh default variant ...   -- ms/step of bench.py per build variant
WORKLOAD=${WORKLOAD:-c3_4k_20obj}; BATCH=${BATCH:-128}
for v in "$@"; do
  if [ "$v" = "default" ]; then unset B200COMP_LIB; else export B200COMP_LIB=$PWD/image_transformation_b200/_lib/variants/$v.so; fi
  ok=$(timeout 150 python -m pytest tests/test_gpu_parity.py -q -x -k "batch or c3 or fuzz" 2>&1 | tail -1)
  ms=$(timeout 90 python bench.py --workload $WORKLOAD --batch $BATCH --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], 'smem', d['config']['smem_bytes_per_cta'])")
  echo "== $v [$WORKLOAD x $BATCH]: $ms | $ok"
done
