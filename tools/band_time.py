"""Device time of a block-row band of one frame pair on ONE GPU (development aid):
usage: python tools/band_time.py W H B R  -> full frame, halves, quarters."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    W, H, B, R = map(int, sys.argv[1:5])
    cur8, ref8 = me.tiled_frames(W, H)
    cur, ref = torch.from_numpy(cur8).cuda(), torch.from_numpy(ref8).cuda()
    with me.Estimator(W, H, B, R) as est:
        nb, nby = est.num_blocks, est.blocks_y
        o = [torch.zeros((1, nb), dtype=torch.int32, device="cuda") for _ in range(3)]
        st = torch.cuda.current_stream().cuda_stream
        for parts in (1, 2, 4, 8):
            rows = nby // parts
            b0 = (nby - rows) // 2
            for _ in range(3):
                est.search_device(cur, ref, W, W * H, 1, o[0], o[1], o[2], None, st, b0, b0 + rows)
            torch.cuda.synchronize()
            ts = []
            for _ in range(15):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                est.search_device(cur, ref, W, W * H, 1, o[0], o[1], o[2], None, st, b0, b0 + rows)
                e1.record()
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"{W}x{H} B={B} R={R}: {rows} of {nby} block rows: median {np.median(ts):.3f} ms  min {min(ts):.3f} ms")


if __name__ == "__main__":
    main()
