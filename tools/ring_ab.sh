#!/bin/bash
# usage (GPU box): tools/ring_ab.sh  -- parity subset, then ms/step for forced ring depths and the automatic choice
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "batch or c3 or fuzz or properties" 2>&1 | tail -2
for r in 3 4 5 ""; do
  for wl in c3_4k_20obj c5_8k_64obj; do
    B200COMP_RING=$r timeout 200 python bench.py --workload $wl $([ $wl = c3_4k_20obj ] && echo --batch 256) --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ring=$r', '$wl', round(d['ms_per_step'],2), 'tile', round(d['roofline']['kernel_split']['tile_kernel_ms_per_step'],2), 'smem', d['config']['smem_bytes_per_cta'], 'pre', d['config']['preresampled_placements'])"
  done
done
