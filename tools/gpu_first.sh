#!/bin/bash
# first GPU bring-up: stage tests with full diagnostics, then pipeline tests
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q --tb=short -x 2>&1 | tail -60 > gpurun_out/stages.log
tail -30 gpurun_out/stages.log
timeout 1200 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short 2>&1 | tail -80 > gpurun_out/pipeline.log
tail -40 gpurun_out/pipeline.log
