"""Runs the integer-pipe microbenchmarks of libme_b200.so (me_b200_int_peak) and
writes gpurun_out/int_peaks.json -- the measured denominators of the integer
roofline.  Usage on the GPU box: python tools/int_peaks.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import motionestimation_b200 as me  # noqa: E402


def main():
    res = {}
    for which, name in enumerate(me.PEAK_NAMES):
        best, mhz_at = 0.0, 0.0
        for _ in range(3):
            rate, mhz = me.int_peak(which, iters=4000)
            if rate > best:
                best, mhz_at = rate, mhz
        res[name] = {"lane_instr_per_s": best, "sm_mhz": mhz_at,
                     "lane_instr_per_clk_per_sm": best / (mhz_at * 1e6) / 148 if mhz_at else None}
        print(name, res[name], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/int_peaks.json", "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
