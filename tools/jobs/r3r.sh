#!/bin/bash
# GPU job: stages as a work queue (re-arm when the last CHUNK is finished) vs the previous ring (exp_v1)
out=gpurun_out/r3r; mkdir -p $out
(timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or random_differential" 2>&1 | tail -4) | tee $out/tests_quick.log
{
for lib in "" motionestimation_b200/libme_b200_exp_v1.so; do
  echo "== library: ${lib:-default (work queue)}"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4" "1920 1080 8 12 16" "1920 1080 16 32 16" "1920 1080 16 8 16" "1920 1080 16 12 16" "1920 1080 16 16 16" "3840 2160 16 32 4" "1920 1080 16 64 8" "1920 1080 16 32 1"; do
    ME_B200_LIBRARY=$lib timeout 120 python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-200
  done
done
} | tee $out/ring.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) | tee $out/tests.log
(timeout 300 python tools/fuzz_parity.py 300 31 mse 2>&1 | tail -3) | tee $out/fuzz.log
