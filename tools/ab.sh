#!/bin/bash
# usage (on the GPU box): tools/ab.sh [variant ...]   -- "" = the default library.  Prints ms/step of bench.py --batch 128
for v in "$@"; do
  if [ "$v" = "default" ]; then unset B200COMP_LIB; else export B200COMP_LIB=$PWD/image_transformation_b200/_lib/variants/$v.so; fi
  ok=$(timeout 150 python -m pytest tests/test_gpu_parity.py -q -x -k "batch" 2>&1 | tail -1)
  ms=$(timeout 90 python bench.py --batch 128 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "import sys,json; print(json.loads(sys.stdin.read())['ms_per_step'])")
  echo "== $v : $ms ms/128 canvases | $ok"
done
