#!/bin/bash
out=gpurun_out/r3c; mkdir -p $out
./tools/fp32x2_probe | tee $out/fp32x2_probe.txt
(timeout 900 python -m pytest tests/test_gpu_ssim.py -m gpu -x -q 2>&1 | tail -8) > $out/tests.log; cat $out/tests.log
