#!/bin/bash
# GPU job: chunk epilogue of the 8x8 kernels -- warp vote for the dx tie-break (default) vs second masked reduction (exp)
out=gpurun_out/r3w; mkdir -p $out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3) | tee $out/tests.log
(timeout 300 python tools/fuzz_parity.py 300 41 mse 2>&1 | tail -3) | tee $out/fuzz.log
{
for lib in "" motionestimation_b200/libme_b200_exp_noballot.so; do
  echo "== library: ${lib:-default (vote)}"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4" "1920 1080 8 12 16" "1920 1080 16 32 16"; do
    ME_B200_LIBRARY=$lib timeout 120 python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-170
  done
done
} | tee $out/epilogue.txt
