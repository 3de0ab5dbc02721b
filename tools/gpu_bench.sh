#!/bin/bash
# bench + ncu launch list + one full capture of the top kernel (B200_PROFILING.md recipe)
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
if [ "$1" = "ncu" ]; then
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:onesweep_pass_kernel -s 40 -c 3 \
    -o gpurun_out/prof_onesweep python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g > gpurun_out/ncu2.log 2>&1
echo "full capture rc=$?"
fi
