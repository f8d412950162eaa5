#!/bin/bash
# usage (GPU box): tools/e2e_sweep.sh chunk ...   -- end-to-end MP/s of bench.py for sub-chunk sizes, 3 runs each
for c in "$@"; do
  for r in 1 2 3; do
    timeout 200 python bench.py --e2e-chunk $c --steps 2 --batch 64 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk $c run $r: e2e %.0f MP/s  %.2f ms/step; solid %.0f' % (d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['solid_canvas_variant']['value']))"
  done
done
