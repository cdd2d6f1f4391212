"""Sweep of the drop-in call's host-side knobs (worker threads, bands, reference-first) on one geometry.
usage: python tools/dropin_sweep.py [W H B R]"""
import itertools
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    W, H, B, R = (int(a) for a in sys.argv[1:5]) if len(sys.argv) > 4 else (1920, 1080, 16, 32)
    cur8, ref8 = me.tiled_frames(W, H)
    cur, ref = cur8.astype(np.int32).ravel(), ref8.astype(np.int32).ravel()
    lib = me.load_library()
    for threads, bands, rf in itertools.product((6, 10, 14), (4, 8), (1, 0)):
        os.environ["ME_B200_PACK_THREADS"] = str(threads)
        os.environ["ME_B200_DROPIN_BANDS"] = str(bands)
        os.environ["ME_B200_DROPIN_REF_FIRST"] = str(rf)
        lib.me_b200_release_cached()
        pf = me.create_prediction_frame(cur, W, H, B)
        for _ in range(5):
            me.search_prediction_frame(pf, ref, R)
        ts = []
        for _ in range(60):
            t0 = time.perf_counter()
            me.search_prediction_frame(pf, ref, R)
            ts.append(time.perf_counter() - t0)
        ts = np.array(ts) * 1e3
        print(f"{W}x{H} B={B} R={R} threads={threads} bands={bands} ref_first={rf}: mean {ts.mean():.3f} ms  median {np.median(ts):.3f}  "
              f"min {ts.min():.3f}", flush=True)
    lib.me_b200_release_cached()


if __name__ == "__main__":
    main()
