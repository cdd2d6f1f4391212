#!/bin/bash
# GPU job: 8x8 blocks with one unrolled copy per period kind (no ramp branches) -- A/B and parity
out=gpurun_out/r3l; mkdir -p $out
{
for pair in 0 1 ""; do
  echo "== ME_B200_PAIR=${pair:-default}"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4" "1920 1080 8 12 16" "1920 1080 8 8 16" "1920 1080 8 16 16"; do
    if [ -n "$pair" ]; then export ME_B200_PAIR=$pair; else unset ME_B200_PAIR; fi
    python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-200
  done
done
unset ME_B200_PAIR
echo "== headline"; python tools/quick_bench.py 1920 1080 16 32 16 2>&1 | grep median | cut -c1-200
python tools/quick_bench.py 3840 2160 16 32 4 2>&1 | grep median | cut -c1-200
} | tee $out/split.txt
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -4) > $out/tests.log; cat $out/tests.log
(timeout 600 python tools/fuzz_parity.py 200 21 mse 2>&1 | tail -4) > $out/fuzz.log; cat $out/fuzz.log
