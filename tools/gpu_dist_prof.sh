#!/bin/bash
# kernel-level launch list of the distributed path at world 1 ($1 = size)
S=${1:-268435456}
mkdir -p gpurun_out
python tools/bench_block.py --size $S --reps 1 > gpurun_out/dist_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/dist_launches.csv \
    python tools/bench_block.py --size $S --reps 1 > gpurun_out/dist_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/dist_plain.log
