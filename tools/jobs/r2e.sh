#!/bin/bash
# GPU job: TMA-fed stream kernel -- tests, timing, stripe sweep, ncu
out=gpurun_out/r2e; mkdir -p $out
(python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "small_span or random_differential" 2>&1 | tail -30) > $out/tests.log
tail -3 $out/tests.log
(for r in 1 2 3 4; do python tools/quick_bench.py 1920 1080 16 $r 64; done
 for n in 5 7 9 10 12 17; do echo stripes=$n; ME_B200_STREAM_STRIPES=$n python tools/quick_bench.py 1920 1080 16 2 64; done
 for n in 5 6 7 9 12; do echo stripes=$n; ME_B200_STREAM_STRIPES=$n python tools/quick_bench.py 1920 1080 16 1 64; done
 python tools/quick_bench.py 3840 2160 16 2 32; python tools/quick_bench.py 3840 2160 16 1 32
 python tools/quick_bench.py 1920 1080 16 32 64 ) > $out/quick.log 2>&1
cat $out/quick.log
NCU="ncu --set full --clock-control none --import-source on"
python tools/quick_bench.py 1920 1080 16 2 64 > $out/plain_r2.log 2>&1 && $NCU -k regex:stream_search -s 3 -c 1 -o $out/prof_stream_r2 python tools/quick_bench.py 1920 1080 16 2 64 > $out/ncu_r2.log 2>&1
python tools/quick_bench.py 1920 1080 16 1 64 > $out/plain_r1.log 2>&1 && $NCU -k regex:stream_search -s 3 -c 1 -o $out/prof_stream_r1 python tools/quick_bench.py 1920 1080 16 1 64 > $out/ncu_r1.log 2>&1
