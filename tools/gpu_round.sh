#!/bin/bash
# usage (on the GPU box): tools/gpu_round.sh [batch]  -- smoke, GPU parity tests, short bench; logs under gpurun_out/
B=${1:-128}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/gputest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/gputest.log
timeout 300 python bench.py --batch $B --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"; cat gpurun_out/bench_quick.json; tail -3 gpurun_out/bench_quick.err
