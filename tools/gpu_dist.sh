#!/bin/bash
# distributed single block on an N-GPU lease: parity tests, then timings.  $1 = N, $2 = sizes (space separated)
N=${1:-2}; SIZES=${2:-"268435456 1073741824"}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q --tb=short 2>&1 | tail -15 | tee gpurun_out/dist_test_n$N.log
for S in $SIZES; do
  for G in 1 2 4 8; do
    if [ $G -le $N ]; then
      if [ $G -eq 1 ]; then
        BZAP_DIST_TIMING=$TIMING timeout 600 python tools/bench_block.py --size $S --single 2>gpurun_out/blk.err | tee -a gpurun_out/block_n$N.jsonl
      else
        timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29571 tools/bench_block.py --size $S 2>gpurun_out/blk.err | tee -a gpurun_out/block_n$N.jsonl
      fi
      tail -3 gpurun_out/blk.err
    fi
  done
done
