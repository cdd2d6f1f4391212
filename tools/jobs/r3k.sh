#!/bin/bash
# GPU job: headline geometry, 16 pairs -- strips per item / parts sweep for FORM 3
out=gpurun_out/r3k; mkdir -p $out
{
echo "== default"; ME_B200_VERBOSE=1 python tools/quick_bench.py 1920 1080 16 32 16 2>&1 | grep "tiled<\|median" | tail -2 | cut -c1-200
for ns in 8 9 10 11; do for parts in 1 2 3 4; do
  echo "== ns=$ns parts=$parts"; ME_B200_NS=$ns ME_B200_PARTS=$parts python tools/quick_bench.py 1920 1080 16 32 16 2>&1 | grep "median" | cut -c1-140
done; done
for parts in 1 2 3; do echo "== 4K parts=$parts"; ME_B200_PARTS=$parts ME_B200_VERBOSE=1 python tools/quick_bench.py 3840 2160 16 32 4 2>&1 | grep "median\|tiled<" | tail -2 | cut -c1-160; done
for parts in 1 2 3 4; do echo "== 1080p +-64 parts=$parts"; ME_B200_PARTS=$parts ME_B200_VERBOSE=1 python tools/quick_bench.py 1920 1080 16 64 8 2>&1 | grep "median\|tiled<" | tail -2 | cut -c1-160; done
} | tee $out/parts.txt
