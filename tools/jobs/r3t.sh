#!/bin/bash
# GPU job: item shapes for 8x8 under the stage queue
out=gpurun_out/r3t; mkdir -p $out
{
for ns in 0 5 7 9 11 14; do for parts in 0 2; do
  echo "== ns=$ns parts=$parts"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4"; do
    if [ $ns = 0 ]; then unset ME_B200_NS; else export ME_B200_NS=$ns; fi
    if [ $parts = 0 ]; then unset ME_B200_PARTS; else export ME_B200_PARTS=$parts; fi
    timeout 120 python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-170
  done
done; done
} | tee $out/items8.txt
