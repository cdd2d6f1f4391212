#!/bin/bash
out=gpurun_out/r3i; mkdir -p $out
nproc
python tools/dropin_sweep.py 1920 1080 16 32 2>&1 | tee $out/sweep_1080p.txt
