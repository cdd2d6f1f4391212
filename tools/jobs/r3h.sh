#!/bin/bash
# GPU job: drop-in call with the reference frame first (FORM 3 on the arriving current frame)
out=gpurun_out/r3h; mkdir -p $out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "drop_in or dropin" 2>&1 | tail -6) > $out/tests.log; cat $out/tests.log
for rf in 1 0; do
  for w in 1080p_16x16_pm32 4k_16x16_pm32 4k_8x8_pm12; do
    ME_B200_DROPIN_REF_FIRST=$rf python bench.py --workload $w --no-cpu-baseline --no-post --sustained-s 0 --no-parity-check --no-band-split --steps 3 > $out/${w}_rf$rf.json 2> $out/${w}_rf$rf.err
    python - <<PY
import json
for l in open("$out/${w}_rf$rf.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$w ref_first=$rf dropin", d.get("e2e_dropin"))
PY
  done
done
ME_B200_TRACE=1 python - <<'PY' 2>&1 | tail -6
import numpy as np, sys
sys.path.insert(0, ".")
import motionestimation_b200 as me
W, H, B, R = 1920, 1080, 16, 32
c8, r8 = me.tiled_frames(W, H, 2, 1)
cur, ref = c8.astype(np.int32).ravel(), r8.astype(np.int32).ravel()
pf = me.create_prediction_frame(cur, W, H, B)
for _ in range(6):
    me.search_prediction_frame(pf, ref, R)
PY
