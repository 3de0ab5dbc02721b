#!/bin/bash
# first GPU bring-up: stage tests with full diagnostics, then pipeline tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
for t in test_mtf_and_inverse test_hist_and_first_appearance test_huffman_encode_decode test_bwt_small test_ibwt; do
  timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -q --tb=short -k $t 2>&1 | tail -40 > gpurun_out/stage_$t.log
  echo "== $t: $(tail -1 gpurun_out/stage_$t.log)"
done
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/stages.log
tail -5 gpurun_out/stages.log
timeout 1500 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short 2>&1 | tail -150 > gpurun_out/pipeline.log
tail -5 gpurun_out/pipeline.log
