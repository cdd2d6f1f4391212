#!/bin/bash
# GPU job: single-pair launches -- strips per item / parts sweep (is the cost model's choice the best one?)
out=gpurun_out/r3j; mkdir -p $out
{
echo "== default"; ME_B200_VERBOSE=1 python tools/quick_bench.py 1920 1080 16 32 1 2>&1 | grep "tiled<\|median" | tail -2 | cut -c1-200
for ns in 2 3 4 5 6 7 9 11; do for parts in 1 2; do
  echo "== ns=$ns parts=$parts"; ME_B200_NS=$ns ME_B200_PARTS=$parts python tools/quick_bench.py 1920 1080 16 32 1 2>&1 | grep "median" | cut -c1-140
done; done
echo "== 4K 16x16 +-32 default"; ME_B200_VERBOSE=1 python tools/quick_bench.py 3840 2160 16 32 1 2>&1 | grep "tiled<\|median" | tail -2 | cut -c1-200
for ns in 3 5 7 9; do echo "== 4K ns=$ns"; ME_B200_NS=$ns python tools/quick_bench.py 3840 2160 16 32 1 2>&1 | grep "median" | cut -c1-140; done
} | tee $out/single_pair.txt
