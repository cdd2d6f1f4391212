"""Regenerates tests/golden/golden_ssim.json + fields_ssim.npz from the UNMODIFIED reference
SSIM search (oracle/_ref/libme_ref_ssim.so = src/cpu/main_ssim.c + src/common/ssim.c compiled
where they lie under /root/reference by `make -C oracle ref`).
Run in the build container only:   python tests/golden/make_golden_ssim.py
The GPU box has no /root/reference; tests there read the committed files."""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_binding import RefSsim, Oracle  # noqa: E402
from cases import SSIM_CASES, make_frames  # noqa: E402


def main():
    ref, orc = RefSsim(), Oracle()
    meta, fields = {}, {}
    for name, gen, args, B, R in SSIM_CASES:
        cur, rf = make_frames(gen, args)
        r = ref.search(cur, rf, B, R)
        o = orc.search_ssim(cur, rf, B, R)
        for k in ("mvx", "mvy", "ssd"):
            assert np.array_equal(r[k], o[k]), (name, k)
        assert np.array_equal(r["score"].view(np.uint32), o["score"].view(np.uint32)), name
        out5, orig, comp = ref.output5(cur, rf, B, r["mvx"], r["mvy"])
        meta[name] = {
            "gen": gen, "args": list(args), "B": B, "R": R, "W": int(cur.shape[1]), "H": int(cur.shape[0]),
            "blocks": int(len(r)),
            "yuv_md5": hashlib.md5(out5.tobytes()).hexdigest(),
            # the line main_ssim.c:95 prints
            "scores_line": "Original Score: %.4f, Compensated Score: %.4f" % (orig, comp),
            "not_found": int(np.count_nonzero(r["ssd"] == 0)),
            "nonzero_mv": int(np.count_nonzero((r["mvx"] != 0) | (r["mvy"] != 0))),
            "cur_md5": hashlib.md5(cur.tobytes()).hexdigest(),
            "ref_md5": hashlib.md5(rf.tobytes()).hexdigest(),
        }
        fields[name + "/mvx"] = r["mvx"].astype(np.int16)
        fields[name + "/mvy"] = r["mvy"].astype(np.int16)
        fields[name + "/score_bits"] = r["score"].view(np.uint32)
        fields[name + "/found"] = r["ssd"].astype(np.uint8)
        print(name, meta[name]["scores_line"], meta[name]["yuv_md5"], "not found:", meta[name]["not_found"])
    with open(os.path.join(HERE, "golden_ssim.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "fields_ssim.npz"), **fields)


if __name__ == "__main__":
    main()
