#!/bin/bash
out=gpurun_out/r3s; mkdir -p $out
{
for lib in "" motionestimation_b200/libme_b200_exp_v1.so; do
  echo "== library: ${lib:-default (work queue, backoff)}"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "1920 1080 16 32 16" "1920 1080 16 8 16" "1920 1080 16 12 16" "3840 2160 16 32 4" "1920 1080 16 64 8" "1920 1080 16 32 1"; do
    ME_B200_LIBRARY=$lib timeout 120 python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-200
  done
done
} | tee $out/ring.txt
