#!/bin/bash
# rebuilds radix_sort.o with each variant's -D flags on the GPU box and times the BWT
mkdir -p gpurun_out
cd bwt_mtf_huffman_compressor_b200/csrc
: > ../../gpurun_out/sweep.log
while read -r tag flags; do
  [ -z "$tag" ] && continue
  rm -f build/*.o
  make -j16 EXTRA="$flags" > /dev/null 2>&1 || { echo "$tag build failed" >> ../../gpurun_out/sweep.log; continue; }
  (cd ../.. && timeout 300 python tools/bwt_time.py "$tag" $SWEEP_MORE >> gpurun_out/sweep.log 2>&1)
done < ../../tools/sweep_variants.txt
rm -f build/*.o; make -j16 > /dev/null 2>&1
cat ../../gpurun_out/sweep.log
