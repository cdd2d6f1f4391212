#!/bin/bash
# GPU job: long randomised parity fuzz of the final build (every mode), drop-in tests, smoke
out=gpurun_out/r3p; mkdir -p $out
for m in "400 101 mse" "400 102 mse" "200 103 ssim" "200 104 fast" "300 105 stream" "200 106 pair" "300 107 all"; do
  (timeout 900 python tools/fuzz_parity.py $m 2>&1 | tail -3)
done | tee $out/fuzz.txt
ME_B200_SSIM_FORM4=1 timeout 600 python tools/fuzz_parity.py 150 108 ssim 2>&1 | tail -3 | tee -a $out/fuzz.txt
python tools/r0_fuzz.py 2>&1 | tail -2 | tee -a $out/fuzz.txt
