#!/bin/bash
# full ncu capture of kernels matching $1 (regex), skipping $2 matching launches, capturing $3
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g --no-calgary > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:$1" -s ${2:-4} -c ${3:-3} \
    -o gpurun_out/prof_$4 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g --no-calgary > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
