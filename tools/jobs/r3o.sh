#!/bin/bash
# GPU job: ring of up to 8 stages (default lib) vs 4 (exp_st4), after the per-period-kind copies
out=gpurun_out/r3o; mkdir -p $out
{
for lib in "" motionestimation_b200/libme_b200_exp_st4.so; do
  echo "== library: ${lib:-default (8 stages)}"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4" "1920 1080 8 12 16" "1920 1080 16 32 16" "1920 1080 16 8 16" "1920 1080 16 16 16" "3840 2160 16 32 4"; do
    ME_B200_LIBRARY=$lib python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-200
  done
done
} | tee $out/stages.txt
