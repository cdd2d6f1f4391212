"""Quick device-resident timing of one geometry (development aid, not the bench contract).
usage: python tools/quick_bench.py W H B R [npairs] [kernel] [reps] [cost: 0 mse | 1 ssim] [search: 0 full | 1 tss | 2 diamond]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    W, H, B, R = map(int, sys.argv[1:5])
    npairs = int(sys.argv[5]) if len(sys.argv) > 5 else 1
    kernel = int(sys.argv[6]) if len(sys.argv) > 6 else 0
    reps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
    cost = int(sys.argv[8]) if len(sys.argv) > 8 else 0
    search = int(sys.argv[9]) if len(sys.argv) > 9 else 0
    cur8, ref8 = me.tiled_frames(W, H)
    pitch = (W + 15) & ~15
    cur = torch.zeros((npairs, H, pitch), dtype=torch.uint8, device="cuda")
    ref = torch.zeros_like(cur)
    cur[:, :, :W] = torch.from_numpy(cur8).cuda()
    ref[:, :, :W] = torch.from_numpy(ref8).cuda()
    with me.Estimator(W, H, B, R, max_pairs=npairs, kernel=kernel, cost=cost, search=search) as est:
        nb = est.num_blocks
        mvx = torch.zeros((npairs, nb), dtype=torch.int32, device="cuda")
        mvy = torch.zeros_like(mvx)
        ssd = torch.zeros_like(mvx)
        sc = torch.zeros((npairs, nb), dtype=torch.float32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            est.search_device(cur, ref, pitch, H * pitch, npairs, mvx, mvy, ssd, sc, st)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            est.search_device(cur, ref, pitch, H * pitch, npairs, mvx, mvy, ssd, sc, st)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        pc = est.pixel_compares * npairs
        if search:
            ev0 = est.fast_evaluations
            est.search_device(cur, ref, pitch, H * pitch, npairs, mvx, mvy, ssd, sc, st)
            ev = est.fast_evaluations - ev0
            print(f"{W}x{H} B={B} R={R} pairs={npairs} search={search}: median {ms:.4f} ms  {npairs / ms * 1e3:.1f} frames/s  "
                  f"{nb * npairs / ms * 1e3 / 1e6:.2f} Mblocks/s  {ev / ms / 1e6:.2f} G candidate evaluations/s "
                  f"({ev / (nb * npairs):.1f} per block)  {ev * B * B / ms / 1e9:.3f} Tpc/s")
            return
        print(f"cost={cost} {W}x{H} B={B} R={R} pairs={npairs} kernel={est.kernel_in_use}: median {ms:.4f} ms  min {min(ts):.4f} ms  "
              f"{npairs / ms * 1e3:.1f} frames/s  {nb * npairs / ms * 1e3 / 1e6:.2f} Mblocks/s  {pc / ms / 1e9:.2f} Tpc/s "
              f"({pc / ms / 1e9 / 73.17 * 100:.1f}% of 73.17 Tpc/s pair peak)")


if __name__ == "__main__":
    main()
