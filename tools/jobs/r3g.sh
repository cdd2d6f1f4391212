#!/bin/bash
# GPU job: ring depth 4 vs 8 at medium spans (experiment library libme_b200_exp_st8.so)
out=gpurun_out/r3g; mkdir -p $out
for lib in "" motionestimation_b200/libme_b200_exp_st8.so; do
  echo "== library: ${lib:-default}"
  for g in "1920 1080 16 8 16" "1920 1080 16 12 16" "1920 1080 16 16 16" "1920 1080 16 32 16" "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4" "1920 1080 16 5 16"; do
    ME_B200_LIBRARY=$lib ME_B200_VERBOSE=1 python tools/quick_bench.py $g 2>&1 | grep -v "^$" | tail -3 | cut -c1-230
  done
done 2>&1 | tee $out/stages.txt
