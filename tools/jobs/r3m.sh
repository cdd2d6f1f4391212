#!/bin/bash
# GPU job: FORM 3 with split period copies (experiment libraries) vs default
out=gpurun_out/r3m; mkdir -p $out
{
for lib in "" motionestimation_b200/libme_b200_exp_s16_1.so motionestimation_b200/libme_b200_exp_s16_2.so; do
  echo "== library: ${lib:-default}"
  for g in "1920 1080 16 32 16" "3840 2160 16 32 4" "1920 1080 16 64 8" "1920 1080 16 16 16" "1920 1080 16 8 16"; do
    ME_B200_LIBRARY=$lib python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-200
  done
done
} | tee $out/split16.txt
