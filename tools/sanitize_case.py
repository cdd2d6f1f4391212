"""Small tuned-kernel invocation for compute-sanitizer (memcheck / racecheck / initcheck):
Foreman 16x16 +-32 and 8x8 +-12, both energy-table and on-the-fly formulations."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402

cur, ref = me.foreman(2), me.foreman(1)
for form in ("2", "1"):
    os.environ["ME_B200_FORM"] = form
    for (B, R) in ((16, 32), (8, 12)):
        with me.Estimator(352, 288, B, R, max_pairs=2) as est:
            out = est.search_u8(np.stack([cur, ref]), np.stack([ref, ref]))
            assert est.kernel_in_use == me.ME_KERNEL_TILED
            assert not out["ssd"][1].any()
# the SSIM cost, on its own streaming kernel and as the tiled kernel's FORM 4
os.environ.pop("ME_B200_FORM", None)
for f4 in ("0", "1"):
    os.environ["ME_B200_SSIM_FORM4"] = f4
    with me.Estimator(352, 288, 16, 12, max_pairs=2, cost=me.ME_COST_SSIM) as est:
        out = est.search_u8(np.stack([cur, ref]), np.stack([ref, ref]))
        assert (out["mvx"][1] == 0).all() and (out["mvy"][1] == 0).all()
os.environ.pop("ME_B200_SSIM_FORM4", None)
# small spans: the TMA-fed streaming kernel
with me.Estimator(352, 288, 16, 2, max_pairs=2) as est:
    out = est.search_u8(np.stack([cur, ref]), np.stack([ref, ref]))
    assert not out["ssd"][1].any()
print("sanitize_case ok")
