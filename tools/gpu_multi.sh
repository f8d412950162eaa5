#!/bin/bash
# usage (on an N-GPU box): tools/gpu_multi.sh N  -- bench lines of C3 (default), C4 and C5 at N GPUs under torchrun
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
numactl -H >> gpurun_out/topo_n$N.txt 2>&1 || lscpu | grep -i numa >> gpurun_out/topo_n$N.txt
timeout 600 $TR bench.py --gpus $N > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err; echo "c3 rc=$?"; cat gpurun_out/bench_r2_n$N.json
[ -z "$SKIP_NOPIN" ] && timeout 600 $TR bench.py --gpus $N --no-numa-pin --no-fresh-plan --steps 3 > gpurun_out/bench_r2_n${N}_nopin.json 2> gpurun_out/bench_r2_n${N}_nopin.err; echo "c3 nopin rc=$?"; cat gpurun_out/bench_r2_n${N}_nopin.json
timeout 600 $TR bench.py --gpus $N --workload c4_aspect_sweep --batch 512 --steps 5 --no-fresh-plan > gpurun_out/bench_r2_n${N}_c4.json 2> gpurun_out/bench_r2_n${N}_c4.err; echo "c4 rc=$?"; cat gpurun_out/bench_r2_n${N}_c4.json
timeout 600 $TR bench.py --gpus $N --workload c5_8k_64obj --steps 5 --no-fresh-plan > gpurun_out/bench_r2_n${N}_c5.json 2> gpurun_out/bench_r2_n${N}_c5.err; echo "c5 rc=$?"; cat gpurun_out/bench_r2_n${N}_c5.json
