#!/bin/bash
# GPU job: SSIM cost on the tiled kernel (FORM 4): parity tests, fuzz, bench A/B against the streaming kernel
out=gpurun_out/r3a; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_ssim.py -m gpu -x -q 2>&1 | tail -15) > $out/tests.log; cat $out/tests.log
(timeout 600 python tools/fuzz_parity.py 150 11 ssim 2>&1 | tail -8) > $out/fuzz.log; cat $out/fuzz.log
for w in ssim_1080p_16x16_pm32 ssim_4k_16x16_pm7; do
  ME_B200_VERBOSE=1 python bench.py --workload $w --no-cpu-baseline > $out/$w.json 2> $out/$w.err
  ME_B200_SSIM_FORM4=0 python bench.py --workload $w --no-cpu-baseline --no-parity-check > $out/${w}_old.json 2> $out/${w}_old.err
  grep -m2 "me_b200\] tiled" $out/$w.err
  python - <<PY
import json
for f in ("$out/$w.json", "$out/${w}_old.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l); print(f, "value", round(d["value"], 1), "frac", round(d["roofline"]["frac"], 4), "parity", d.get("parity_checked"), "launches", d.get("gpu_launches"))
PY
done
