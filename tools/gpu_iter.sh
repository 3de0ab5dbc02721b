#!/bin/bash
# iteration loop on the GPU: parity first, then the bench line; optional ncu capture of one kernel ($1 = regex)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/stages.log
tail -4 gpurun_out/stages.log
timeout 1500 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --tb=short -x 2>&1 | tail -25 > gpurun_out/pipeline.log
tail -4 gpurun_out/pipeline.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ -n "$1" ]; then
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-4} -c ${3:-3} \
    -o gpurun_out/prof_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
fi
