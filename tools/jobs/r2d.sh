#!/bin/bash
# GPU job: small-span kernel tests after the fix, ncu captures of the stream kernel (+-2, +-1) and of the
# tuned kernel on 4K 8x8 +-12, drop-in phase trace
out=gpurun_out/r2d; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_span or random_differential or postprocess or drop_in or full_size" 2>&1 | tail -30) > $out/tests.log
tail -3 $out/tests.log
ME_B200_TRACE=1 python tools/dropin_latency.py > $out/dropin.log 2> $out/dropin_trace.log
tail -4 $out/dropin_trace.log
NCU="ncu --set full --clock-control none --import-source on"
python tools/quick_bench.py 1920 1080 16 2 64 > $out/plain_r2.log 2>&1 && $NCU -k regex:stream_search -s 3 -c 1 -o $out/prof_stream_r2 python tools/quick_bench.py 1920 1080 16 2 64 > $out/ncu_r2.log 2>&1
python tools/quick_bench.py 1920 1080 16 1 64 > $out/plain_r1.log 2>&1 && $NCU -k regex:stream_search -s 3 -c 1 -o $out/prof_stream_r1 python tools/quick_bench.py 1920 1080 16 1 64 > $out/ncu_r1.log 2>&1
python tools/quick_bench.py 3840 2160 8 12 4 > $out/plain_8x8.log 2>&1 && $NCU -k regex:tiled_search -s 3 -c 1 -o $out/prof_tiled_8x8_pm12 python tools/quick_bench.py 3840 2160 8 12 4 > $out/ncu_8x8.log 2>&1
cat $out/plain_r2.log $out/plain_r1.log $out/plain_8x8.log
ls -la $out
