#!/bin/bash
# GPU job: pair kernel with 16 warps; drop-in band sweep
out=gpurun_out/r2g; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "random_differential or full_size or ingest" 2>&1 | tail -30) > $out/tests.log
tail -3 $out/tests.log
(for cfg in "3840 2160 8 12 16" "3840 2160 8 32 16" "352 288 8 12 512" "1920 1080 8 12 32" "3840 2160 8 8 16" "3840 2160 8 16 16" "3840 2160 8 64 8"; do
   python tools/quick_bench.py $cfg; ME_B200_PAIR=0 python tools/quick_bench.py $cfg; done) > $out/quick.log 2>&1
cat $out/quick.log
(for b in 2 3 4 5; do echo bands=$b; ME_B200_DROPIN_BANDS=$b python tools/dropin_latency.py; done) > $out/dropin.log 2>&1; cat $out/dropin.log
