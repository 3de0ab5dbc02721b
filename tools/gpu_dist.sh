#!/bin/bash
# distributed single block on an N-GPU lease: parity tests, then timings.
#   $1 = N   $2 = sizes (space separated)   $3 = world sizes to time   $4 = pytest -k filter
N=${1:-2}; SIZES=${2:-"1073741824"}; WORLDS=${3:-"1 2 4 8"}; FILT=${4:-"world"}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q --tb=short -k "$FILT" 2>&1 | tail -15 | tee gpurun_out/dist_test_n$N.log
for S in $SIZES; do
  for G in $WORLDS; do
    if [ $G -le $N ]; then
      if [ $G -eq 1 ]; then
        BZAP_DIST_TIMING=$TIMING timeout 600 python tools/bench_block.py --size $S --single 2>gpurun_out/blk.err | tee -a gpurun_out/block_n$N.jsonl
      else
        BZAP_DIST_TIMING=$TIMING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29571 tools/bench_block.py --size $S 2>gpurun_out/blk.err | tee -a gpurun_out/block_n$N.jsonl
      fi
      grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/blk.err | tail -${TAILN:-3}
    fi
  done
done
# optional NCCL tuning experiment: NCCLX="VAR=val VAR=val" re-times the largest world with that environment
if [ -n "$NCCLX" ]; then
  for S in $SIZES; do G=$N
    env $NCCLX BZAP_DIST_TIMING=$TIMING timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29572 tools/bench_block.py --size $S 2>gpurun_out/blk2.err | tee -a gpurun_out/block_n$N.jsonl
    grep -v "OMP_NUM_THREADS\|^\*\*\*\*" gpurun_out/blk2.err | tail -${TAILN:-3}
  done
fi
