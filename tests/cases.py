"""Shared definition of the golden cases (used by tests and by
tests/golden/make_golden.py, which pins them against the unmodified reference)."""
import json
import os

import numpy as np

from motionestimation_b200 import frames

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, generator, args, B, R
    ("foreman_yf4_yf1_8_12", "foreman", (4, 1), 8, 12),     # reference default run (run.sh:4), PSNR 31.816000
    ("foreman_yf1_yf4_8_12", "foreman", (1, 4), 8, 12),     # results/cpu/foreman/8_12.txt, PSNR 31.750712
    ("foreman_yf4_yf1_4_15", "foreman", (4, 1), 4, 15),     # results/cpu/foreman/output_4_15.yuv
    ("foreman_yf4_yf1_4_7", "foreman", (4, 1), 4, 7),       # results/cpu/foreman/output_4_7.yuv
    ("foreman_yf2_yf1_8_12", "foreman", (2, 1), 8, 12),     # BASELINE config 1
    ("foreman_yf2_yf1_8_32", "foreman", (2, 1), 8, 32),
    ("foreman_yf2_yf1_16_32", "foreman", (2, 1), 16, 32),
    ("foreman_yf2_yf1_16_64", "foreman", (2, 1), 16, 64),
    ("foreman_yf2_yf1_5_7", "foreman", (2, 1), 5, 7),       # partial edge blocks (352 = 70*5+2)
    ("foreman_yf2_yf1_7_9", "foreman", (2, 1), 7, 9),
    ("foreman_yf2_yf1_32_16", "foreman", (2, 1), 32, 16),   # w*h > 256: float-score path
    ("foreman_yf2_yf1_64_8", "foreman", (2, 1), 64, 8),
    ("constant_352x288_8_12", "constant", (352, 288), 8, 12),  # all-tie
    ("noise_200x120_16_32", "shifted_noise", (200, 120, 99), 16, 32),
    ("noise_96x64_16_64", "shifted_noise", (96, 64, 5), 16, 64),   # frame smaller than the window
    ("random_64x48_8_4", "random", (64, 48, 3), 8, 4),
    ("checker_64x64_32_8", "checker", (64, 64), 32, 8),     # SSD >= 2^24: float rounding path
    ("checker_70x66_64_3", "checker", (70, 66), 64, 3),
    ("far_96x80_32_8", "far", (96, 80, 11), 32, 8),         # every SSD >= 2^24 (float rounding decides)
    ("far_100x70_24_5", "far", (100, 70, 12), 24, 5),
    ("far_64x64_16_8", "far", (64, 64, 13), 16, 8),         # 256 px: still exact
]

# SSIM-cost full search (src/cpu/main_ssim.c + src/common/ssim.c); fixtures made by the unmodified
# reference through tests/golden/make_golden_ssim.py
SSIM_CASES = [
    ("ssim_foreman_yf4_yf1_16_7", "foreman", (4, 1), 16, 7),    # main_ssim.c:41-42 defaults (blk 16, span 7)
    ("ssim_foreman_yf4_yf1_4_15", "foreman", (4, 1), 4, 15),    # src/cpu/run_ssim.sh:4
    ("ssim_foreman_yf2_yf1_8_12", "foreman", (2, 1), 8, 12),
    ("ssim_foreman_yf2_yf1_16_32", "foreman", (2, 1), 16, 32),
    ("ssim_foreman_yf2_yf1_5_7", "foreman", (2, 1), 5, 7),      # partial edge blocks
    ("ssim_foreman_yf2_yf1_7_9", "foreman", (2, 1), 7, 9),
    ("ssim_foreman_yf2_yf1_32_16", "foreman", (2, 1), 32, 16),  # w*h > 258: literal float cross sum
    ("ssim_foreman_yf1_yf2_64_8", "foreman", (1, 2), 64, 8),
    ("ssim_constant_96x64_8_12", "constant", (96, 64), 8, 12),  # every candidate scores exactly 1: first wins
    ("ssim_noise_200x120_16_32", "shifted_noise", (200, 120, 99), 16, 32),
    ("ssim_noise_96x64_16_64", "shifted_noise", (96, 64, 5), 16, 64),
    ("ssim_noise_100x60_8_12", "shifted_noise", (100, 60, 7), 8, 12),   # partial right/bottom blocks, B = 8
    ("ssim_random_64x48_8_4", "random", (64, 48, 3), 8, 4),
    ("ssim_checker_64x64_32_8", "checker", (64, 64), 32, 8),
    ("ssim_far_96x80_32_8", "far", (96, 80, 11), 32, 8),
    ("ssim_far_64x64_16_8", "far", (64, 64, 13), 16, 8),
    ("ssim_inverted_96x80_16_3", "inverted", (96, 80, 21), 16, 3),      # no candidate above 0
    ("ssim_inverted_100x70_8_6", "inverted", (100, 70, 22), 8, 6),
]


def make_frames(gen, args):
    if gen == "foreman":
        return frames.foreman(args[0]), frames.foreman(args[1])
    if gen == "constant":
        return frames.constant_pair(*args)
    if gen == "shifted_noise":
        return frames.shifted_noise_pair(args[0], args[1], seed=args[2])
    if gen == "random":
        return frames.random_pair(args[0], args[1], seed=args[2])
    if gen == "far":
        return frames.far_pair(args[0], args[1], seed=args[2])
    if gen == "checker":
        return frames.checker_pair(*args)
    if gen == "inverted":
        return frames.inverted_pair(args[0], args[1], seed=args[2])
    raise ValueError(gen)



def load_golden():
    with open(os.path.join(HERE, "golden", "golden.json")) as f:
        meta = json.load(f)
    fields = np.load(os.path.join(HERE, "golden", "fields.npz"))
    return meta, fields


def load_golden_ssim():
    with open(os.path.join(HERE, "golden", "golden_ssim.json")) as f:
        meta = json.load(f)
    fields = np.load(os.path.join(HERE, "golden", "fields_ssim.npz"))
    return meta, fields


def case_ids():
    return [c[0] for c in CASES]
