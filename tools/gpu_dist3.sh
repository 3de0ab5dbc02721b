#!/bin/bash
# weak-scaling bench + one distributed 1 GiB block on all GPUs of the box (no single-GPU leg)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}.json 2> gpurun_out/bench_${N}.err
echo "bench N=$N rc=$?"; cut -c1-300 gpurun_out/bench_${N}.json; grep -i "error" gpurun_out/bench_${N}.err | head -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 tools/bench_block.py --size 1073741824 --reps 1 > gpurun_out/block1g_${N}.json 2> gpurun_out/block1g_${N}.err
echo "block1g rc=$?"; cat gpurun_out/block1g_${N}.json; grep -i "error" gpurun_out/block1g_${N}.err | head -5
