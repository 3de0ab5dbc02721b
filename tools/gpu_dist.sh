#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q --tb=short -x 2>&1 | tail -30 > gpurun_out/dist_test.log
tail -6 gpurun_out/dist_test.log
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/bench_block.py --size 268435456 --check > gpurun_out/block_${N}.json 2> gpurun_out/block_${N}.err
echo "block rc=$?"; cat gpurun_out/block_${N}.json; tail -3 gpurun_out/block_${N}.err
if [ "$N" -gt 1 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}.json 2> gpurun_out/bench_${N}.err
echo "bench N=$N rc=$?"; cat gpurun_out/bench_${N}.json | cut -c1-600; tail -3 gpurun_out/bench_${N}.err
fi
