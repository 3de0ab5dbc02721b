#!/bin/bash
# round-end evidence: bench line, launch list of the same workload, full ncu capture of the top kernel
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json
bash tools/gpu_launches.sh
bash tools/gpu_ncu.sh "onesweep_pass_kernel" 140 10 onesweep_final
