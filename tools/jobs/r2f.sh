#!/bin/bash
# GPU job: pair kernel (8x8, two block rows per item) -- parity + timing vs the single-row kernel; drop-in latency
out=gpurun_out/r2f; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_differential or full_size or golden or repeat" 2>&1 | tail -30) > $out/tests.log
tail -3 $out/tests.log
(for cfg in "3840 2160 8 12 16" "3840 2160 8 32 16" "352 288 8 12 512" "1920 1080 8 12 32" "3840 2160 8 8 16" "3840 2160 8 16 16"; do
   python tools/quick_bench.py $cfg; ME_B200_PAIR=0 python tools/quick_bench.py $cfg; done) > $out/quick.log 2>&1
cat $out/quick.log
python tools/dropin_latency.py > $out/dropin.log 2>&1; cat $out/dropin.log
ME_B200_TRACE=1 python tools/dropin_latency.py 2>&1 | grep "bands 4" | tail -3
