"""Latency of the reference drop-in call me_b200_search (int frames + predictionFrame structs,
the replacement of main.c:144-158) and of the blocking u8 call, per frame pair."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    for (W, H, B, R, name) in ((352, 288, 8, 12, "Foreman CIF 8x8 +-12"), (1920, 1080, 16, 32, "1080p 16x16 +-32")):
        if W == 352:
            cur8, ref8 = me.foreman(2), me.foreman(1)
        else:
            cur8, ref8 = me.tiled_frames(W, H)
        cur, ref = cur8.astype(np.int32).ravel(), ref8.astype(np.int32).ravel()
        pf = me.create_prediction_frame(cur, W, H, B)
        me.search_prediction_frame(pf, ref, R)
        n = 50
        t0 = time.perf_counter()
        for _ in range(n):
            me.search_prediction_frame(pf, ref, R)
        t_int = (time.perf_counter() - t0) / n
        with me.Estimator(W, H, B, R) as est:
            est.search_u8(cur8, ref8)
            t0 = time.perf_counter()
            for _ in range(n):
                est.search_u8(cur8, ref8)
            t_u8 = (time.perf_counter() - t0) / n
        print(f"{name}: me_b200_search (int frames, blocks filled) {t_int * 1e3:.3f} ms/call; "
              f"me_b200_search_u8 (pageable u8) {t_u8 * 1e3:.3f} ms/call", flush=True)


if __name__ == "__main__":
    main()
