#!/bin/bash
# GPU job: per-item timeline of the tiled kernel (experiment library with -DME_TRACE_ITEMS)
out=gpurun_out/r3q; mkdir -p $out
export ME_B200_LIBRARY=motionestimation_b200/libme_b200_exp_trace.so ME_B200_TRACE_ITEMS=1 ME_B200_VERBOSE=1
{
for cfg in "3840 2160 8 12 8" "352 288 8 12 256" "1920 1080 16 32 16" "1920 1080 16 8 16" "3840 2160 8 32 4"; do
  echo "== $cfg"; python tools/quick_bench.py $cfg 0 2 2>&1 | grep "item trace\|tiled<\|median" | tail -4 | cut -c1-600
done
echo "== 4K 8x8 +-12 ns=6"; ME_B200_NS=6 python tools/quick_bench.py 3840 2160 8 12 8 0 2 2>&1 | grep "item trace\|tiled<\|median" | tail -3 | cut -c1-600
} | tee $out/trace.txt
