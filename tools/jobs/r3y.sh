#!/bin/bash
out=gpurun_out/r3y; mkdir -p $out
{
for lib in "" motionestimation_b200/libme_b200_exp_ballot16.so; do
  echo "== library: ${lib:-default}"
  for g in "1920 1080 16 32 16" "3840 2160 16 32 4" "1920 1080 16 64 8" "1920 1080 16 8 16" "1920 1080 16 12 16" "1920 1080 16 16 16" "1920 1080 16 32 1"; do
    ME_B200_LIBRARY=$lib timeout 120 python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-170
  done
done
ME_B200_LIBRARY=motionestimation_b200/libme_b200_exp_ballot16.so timeout 300 python tools/fuzz_parity.py 150 51 mse 2>&1 | tail -2
} | tee $out/vote16.txt
