"""Randomised parity fuzz on the GPU: random geometries / spans / frame contents, AUTO kernel (and the
forced formulations) against the oracle.
usage: python tools/fuzz_parity.py [n_cases] [seed] [mode: mse | ssim | fast | all | stream | pair]
(stream: the small-span streaming kernel -- 16x16 blocks, widths that are multiples of 16, spans 1..4, any
height; pair: 8x8 blocks with the energy table and the two-rows-per-item kernel forced)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import motionestimation_b200 as me  # noqa: E402
from oracle_binding import Oracle  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    mode_arg = sys.argv[3] if len(sys.argv) > 3 else "mse"
    rng = np.random.Generator(np.random.PCG64(seed))
    orc = Oracle()
    bad = 0
    for i in range(n):
        B = int(rng.choice([8, 16, 8, 16, 4, 12, 32]))
        R = int(rng.choice([0, 1, 2, 3, 4, 5, 7, 8, 12, 13, 16, 24, 31, 32, 33, 40, 64]))
        W = int(rng.integers(B, 420))
        H = int(rng.integers(B, 300))
        if rng.random() < 0.5:
            W = (W // B) * B
        if rng.random() < 0.3:
            H = (H // B) * B + (B // 2 if rng.random() < 0.5 else 0)
        H = max(H, B)
        if mode_arg == "stream":
            B, R = 16, int(rng.integers(1, 5))
            W = 16 * int(rng.integers(1, 70))
            H = int(rng.integers(16, 300))
        if mode_arg == "pair":
            B, R = 8, int(rng.choice([1, 3, 5, 8, 12, 13, 16, 24, 32, 40]))
            W = 8 * int(rng.integers(2, 60))
            H = int(rng.integers(8 * (2 * ((R + 7) // 8) + 10), 8 * (2 * ((R + 7) // 8) + 40)))
        mode = mode_arg if mode_arg not in ("all", "stream", "pair") else (
            "mse" if mode_arg != "all" else str(rng.choice(["mse", "ssim", "fast"])))
        kind = int(rng.integers(0, 6 if mode == "ssim" else 5))
        if kind == 0:
            cur, ref = me.random_pair(W, H, int(rng.integers(1 << 30)))
        elif kind == 1:
            cur, ref = me.shifted_noise_pair(W, H, seed=int(rng.integers(1 << 30)),
                                             shift=(int(rng.integers(-9, 10)), int(rng.integers(-9, 10))))
        elif kind == 2:
            cur, ref = me.constant_pair(W, H, int(rng.integers(0, 256)))
        elif kind == 3:
            cur, ref = me.checker_pair(W, H, int(rng.integers(1, 5)))
        elif kind == 4:
            cur, ref = me.far_pair(W, H, int(rng.integers(1 << 30)))
        else:
            cur, ref = me.inverted_pair(W, H, seed=int(rng.integers(1 << 30)), period=float(rng.uniform(5, 40)))
        form = str(rng.choice(["", "", "2", "1", "0"]))
        if mode_arg == "pair":
            form = "2"
            os.environ["ME_B200_PAIR"] = "1"
        if form:
            os.environ["ME_B200_FORM"] = form
        else:
            os.environ.pop("ME_B200_FORM", None)
        if mode == "ssim":
            if B == 32 and R > 16:
                R = 16          # keep the CPU side of the fuzz short
            exp = orc.search_ssim(cur, ref, B, R)
            kw = dict(cost=me.ME_COST_SSIM, kernel=int(rng.choice([me.ME_KERNEL_AUTO, me.ME_KERNEL_AUTO, me.ME_KERNEL_GENERIC])))
        elif mode == "fast":
            algo = int(rng.integers(1, 3))
            exp, _ = orc.search_fast(cur, ref, B, R, algo)
            kw = dict(search=algo)
        else:
            exp = orc.search(cur, ref, B, R)
            kw = {}
        try:
            with me.Estimator(W, H, B, R, **kw) as est:
                out = est.search_u8(cur, ref)
                kern = est.kernel_in_use
        except Exception as ex:
            bad += 1
            print(f"ERROR case {i}: mode={mode} W={W} H={H} B={B} R={R} kind={kind} form={form!r} {kw}: {ex}", flush=True)
            continue
        ok = (np.array_equal(out["mvx"][0], exp["mvx"]) and np.array_equal(out["mvy"][0], exp["mvy"]) and
              np.array_equal(out["ssd"][0], exp["ssd"]) and
              np.array_equal(out["score"][0].view(np.uint32), exp["score"].view(np.uint32)))
        if not ok:
            bad += 1
            print(f"MISMATCH case {i}: mode={mode} W={W} H={H} B={B} R={R} kind={kind} form={form!r} kernel={kern} {kw}", flush=True)
    print(f"fuzz[{mode_arg}]: {n} cases, {bad} mismatches (seed {seed})")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
