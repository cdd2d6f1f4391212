#!/bin/bash
# GPU job: FORM 3 (16x16 table with the bias, no sum cur^2 in the tasks) -- parity + A/B timing
out=gpurun_out/r2m; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py tests/test_gpu_peer.py -m gpu -x -q 2>&1 | tail -8) > $out/tests.log; tail -3 $out/tests.log
(for cfg in "1920 1080 16 32 64" "1920 1080 16 64 32" "3840 2160 16 32 16" "1920 1080 16 12 64" "1920 1080 16 8 64" "1920 1080 16 32 1"; do
   python tools/quick_bench.py $cfg; ME_B200_FORM16=2 python tools/quick_bench.py $cfg; done) > $out/quick.log 2>&1
cat $out/quick.log
python tools/fuzz_parity.py 300 21 mse > $out/fuzz.log 2>&1; tail -2 $out/fuzz.log
