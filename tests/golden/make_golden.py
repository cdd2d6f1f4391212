"""Regenerates tests/golden/golden.json + fields.npz from the UNMODIFIED reference
(oracle/_ref/libme_ref.so, built from /root/reference by `make -C oracle ref`).
Run in the build container only:   python tests/golden/make_golden.py
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

from motionestimation_b200 import frames  # noqa: E402
from oracle_binding import Ref, Oracle, field_sha  # noqa: E402

from cases import CASES, make_frames  # noqa: E402


def main():
    ref, orc = Ref(), Oracle()
    meta, fields = {}, {}
    for name, gen, args, B, R in CASES:
        cur, rf = make_frames(gen, args)
        r = ref.search(cur, rf, B, R)
        o = orc.search(cur, rf, B, R)
        assert np.array_equal(r["mvx"], o["mvx"]) and np.array_equal(r["mvy"], o["mvy"]), name
        assert np.array_equal(r["score"].view(np.uint32), o["score"].view(np.uint32)), name
        out5, psnr = ref.output5(cur, rf, B, r["mvx"], r["mvy"])
        meta[name] = {
            "gen": gen, "args": list(args), "B": B, "R": R, "W": int(cur.shape[1]), "H": int(cur.shape[0]),
            "blocks": int(len(r)),
            "yuv_md5": hashlib.md5(out5.tobytes()).hexdigest(),
            "psnr": "%.6f" % psnr,
            "field_sha": field_sha(r["mvx"], r["mvy"], o["ssd"]),
            "nonzero_mv": int(np.count_nonzero((r["mvx"] != 0) | (r["mvy"] != 0))),
            "cur_md5": hashlib.md5(cur.tobytes()).hexdigest(),
            "ref_md5": hashlib.md5(rf.tobytes()).hexdigest(),
        }
        fields[name + "/mvx"] = r["mvx"].astype(np.int16)
        fields[name + "/mvy"] = r["mvy"].astype(np.int16)
        fields[name + "/score_bits"] = r["score"].view(np.uint32)
        fields[name + "/ssd"] = o["ssd"]
        print(name, meta[name]["psnr"], meta[name]["yuv_md5"], meta[name]["field_sha"])
    # the reference's own shipped goldens
    for f in ("output_4_15.yuv", "output_4_7.yuv"):
        p = os.path.join("/root/reference/results/cpu/foreman", f)
        meta["shipped/" + f] = {"md5": hashlib.md5(open(p, "rb").read()).hexdigest()}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "fields.npz"), **fields)


if __name__ == "__main__":
    main()
