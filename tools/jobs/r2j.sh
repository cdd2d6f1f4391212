#!/bin/bash
# GPU job: batch-size sensitivity of the stream kernel (multi-wave launches backfill the staggered finish),
# parity fuzz of the final build
out=gpurun_out/r2j; mkdir -p $out
(for r in 1 2 4; do for np in 64 128 256; do python tools/quick_bench.py 1920 1080 16 $r $np; done; done
 for n in 5 6 7 9; do echo stripes=$n; ME_B200_STREAM_STRIPES=$n python tools/quick_bench.py 1920 1080 16 2 256; done
 for n in 5 6 7 9; do echo stripes=$n; ME_B200_STREAM_STRIPES=$n python tools/quick_bench.py 1920 1080 16 1 256; done) > $out/quick.log 2>&1
cat $out/quick.log
(python tools/fuzz_parity.py 300 11 stream; python tools/fuzz_parity.py 200 12 pair; python tools/fuzz_parity.py 300 13 mse) > $out/fuzz.log 2>&1; tail -5 $out/fuzz.log
