#!/bin/bash
# GPU job: upper bound of dropping sum cur^2 from the 16x16 tasks (experiment build, wrong results); +-0 batch size
out=gpurun_out/r2l; mkdir -p $out
(for i in 1 2; do python tools/quick_bench.py 1920 1080 16 32 64
 ME_B200_LIBRARY=$PWD/motionestimation_b200/libme_b200_exp_nocursq.so python tools/quick_bench.py 1920 1080 16 32 64; done
 python tools/quick_bench.py 3840 2160 16 32 16
 ME_B200_LIBRARY=$PWD/motionestimation_b200/libme_b200_exp_nocursq.so python tools/quick_bench.py 3840 2160 16 32 16
 python tools/quick_bench.py 1920 1080 16 0 64; python tools/quick_bench.py 1920 1080 16 0 256; python tools/quick_bench.py 3840 2160 16 0 64) > $out/quick.log 2>&1
cat $out/quick.log
