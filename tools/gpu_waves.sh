#!/bin/bash
# on the GPU box: parity of the wave-split run + ms/step per wave count (B200COMP_WAVES; 1 = binning and tile kernel one after another)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "batch or c3 or fuzz or properties" 2>&1 | tail -3
for w in 1 4 8 16 ""; do
  for b in 1024 256; do
    ms=$(B200COMP_WAVES=$w timeout 200 python bench.py --batch $b --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan 2>gpurun_out/waves.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])")
    echo "== waves='$w' batch=$b: $ms"
  done
done
for wl in c4_aspect_sweep c5_8k_64obj; do
  for w in 1 ""; do
    ms=$(B200COMP_WAVES=$w timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-fresh-plan 2>gpurun_out/waves.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['frac'], d['gpu_launches'])")
    echo "== $wl waves='$w': $ms"
  done
done
