#!/bin/bash
# GPU job: 8x8 -- chunk index prefetch (default lib) vs none (exp_nopf), and 8 stages with smaller items (exp_st8)
out=gpurun_out/r3n; mkdir -p $out
{
for lib in "" motionestimation_b200/libme_b200_exp_nopf.so; do
  echo "== library: ${lib:-default (prefetch)}"
  for g in "3840 2160 8 12 8" "352 288 8 12 256" "3840 2160 8 32 4" "1920 1080 8 12 16"; do
    ME_B200_LIBRARY=$lib python tools/quick_bench.py $g 2>&1 | grep median | cut -c1-200
  done
done
for ns in 0 6 8 10; do
  echo "== 8 stages, ns=$ns"
  for g in "3840 2160 8 12 8" "352 288 8 12 256"; do
    if [ $ns = 0 ]; then unset ME_B200_NS; else export ME_B200_NS=$ns; fi
    ME_B200_VERBOSE=1 ME_B200_LIBRARY=motionestimation_b200/libme_b200_exp_st8.so python tools/quick_bench.py $g 2>&1 | grep "median\|tiled<" | tail -2 | cut -c1-200
  done
done
} | tee $out/prefetch_stages.txt
