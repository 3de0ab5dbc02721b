#!/bin/bash
# round evidence on a 1-GPU lease: the bench line (with cpu_baseline and the 1 GiB block), the reference arm,
# the ncu launch list of the same workload and full captures of the roofline kernels.  The .ncu-rep files stay
# on the box (/tmp): only their summaries (tools/ncu_summary.py) come back -- gpurun_out/ is capped at 64 MiB.
#   $1 = tag   $2 = "ref" to time the reference arm   $3 = "tests" to run the parity suite first
R=${1:-r2}
mkdir -p gpurun_out
if [ "$3" = "tests" ]; then timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 | tee gpurun_out/${R}_gputests.log; fi
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench.json 2> gpurun_out/${R}_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${R}_bench.json
if [ "$2" = "ref" ]; then
  python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_ref.err; echo "reference rc=$?"
fi
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-block1g --no-calgary > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-block1g --no-calgary > gpurun_out/ncu.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/${R}_launches.csv > gpurun_out/${R}_launches_summary.txt
ncu --set full --clock-control none -k regex:onesweep_pass_kernel -s 40 -c 10 \
    -o /tmp/${R}_prof_onesweep python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g --no-calgary > gpurun_out/ncu2.log 2>&1
echo "onesweep capture rc=$?"
python tools/ncu_summary.py /tmp/${R}_prof_onesweep.ncu-rep > gpurun_out/${R}_onesweep_ncu_summary.txt
ncu --set full --clock-control none -k regex:"ibwt_walk_len_kernel|ibwt_copy_slots|mtf_apply|imtf_apply|imtf_perm|huff_dec_write|huff_encode|bwt_rerank|scatter_u32|bwt_gather" -s 20 -c 20 \
    -o /tmp/${R}_prof_second python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-block1g --no-calgary > gpurun_out/ncu3.log 2>&1
echo "second tier capture rc=$?"
python tools/ncu_summary.py /tmp/${R}_prof_second.ncu-rep > gpurun_out/${R}_second_tier_ncu_summary.txt
du -sh gpurun_out
