"""CPU tests of the checker itself (no GPU): the C restatement under oracle/ is
pinned against (1) the reference's shipped golden outputs, (2) fixtures produced
by the unmodified reference (tests/golden, made by tests/golden/make_golden.py),
(3) the unmodified reference run live when oracle/_ref is present, and (4) its own
literal float-accumulation mode."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import CASES, make_frames, load_golden
from oracle_binding import Oracle, Ref, field_sha, ROOT

META, FIELDS = load_golden()


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_oracle_matches_reference_fixture(orc, case):
    name, gen, args, B, R = case
    cur, ref = make_frames(gen, args)
    m = META[name]
    assert hashlib.md5(cur.tobytes()).hexdigest() == m["cur_md5"], "synthetic generator drifted"
    assert hashlib.md5(ref.tobytes()).hexdigest() == m["ref_md5"]
    o = orc.search(cur, ref, B, R)
    assert len(o) == m["blocks"]
    assert np.array_equal(o["mvx"], FIELDS[name + "/mvx"])
    assert np.array_equal(o["mvy"], FIELDS[name + "/mvy"])
    assert np.array_equal(o["score"].view(np.uint32), FIELDS[name + "/score_bits"])
    assert field_sha(o["mvx"], o["mvy"], o["ssd"]) == m["field_sha"]
    out5, psnr = orc.output5(cur, ref, B, o)
    assert hashlib.md5(out5.tobytes()).hexdigest() == m["yuv_md5"]
    assert "%.6f" % psnr == m["psnr"]


def test_shipped_reference_goldens(orc):
    """results/cpu/foreman/output_4_15.yuv and output_4_7.yuv."""
    assert META["foreman_yf4_yf1_4_15"]["yuv_md5"] == META["shipped/output_4_15.yuv"]["md5"] == \
        "686f3f74e7dc7f2e8321b513eed033e8"
    assert META["foreman_yf4_yf1_4_7"]["yuv_md5"] == META["shipped/output_4_7.yuv"]["md5"] == \
        "a77c2741268fc18e5f73c599fc040172"
    # logged PSNR lines: 2990wx_threadripper_64_cores.txt:10 and 8_12.txt:10
    assert META["foreman_yf4_yf1_8_12"]["psnr"] == "31.816000"
    assert META["foreman_yf1_yf4_8_12"]["psnr"] == "31.750712"


def test_tie_break_known_answer(orc):
    """Two identical constant frames: every candidate ties at 0, the first one in
    visit order wins => MV = (-min(R,x0), -min(R,y0))."""
    from motionestimation_b200 import frames
    cur, ref = frames.constant_pair(352, 288)
    o = orc.search(cur, ref, 8, 12)
    x0, y0, _, _ = frames.block_grid(352, 288, 8)
    assert np.array_equal(o["mvx"], -np.minimum(12, x0))
    assert np.array_equal(o["mvy"], -np.minimum(12, y0))
    assert not o["ssd"].any() and not o["score"].any()
    assert (o["mvx"][0], o["mvx"][1], o["mvx"][2], o["mvy"][45], o["mvx"][45]) == (0, -8, -12, -8, -8)


@pytest.mark.parametrize("B,R,W,H", [(8, 3, 40, 24), (16, 5, 50, 37), (24, 4, 60, 50), (32, 6, 70, 66), (5, 2, 23, 17)])
def test_literal_float_accumulation_agrees(B, R, W, H):
    """The integer-SSD shortcut equals the literal float loop of main.c:19-26,
    including frames whose SSDs exceed 2^24."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from oracle_binding import Oracle; from motionestimation_b200 import frames\n"
        "o = Oracle(); out = []\n"
        "for cur, ref in (frames.far_pair(%d, %d, 3), frames.random_pair(%d, %d, 4)):\n"
        "    r = o.search(cur, ref, %d, %d, nthreads=4)\n"
        "    out.append(np.concatenate([r['mvx'], r['mvy'], r['score'].view(np.int32)]))\n"
        "sys.stdout.write(' '.join(map(str, np.concatenate(out).tolist())))\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), W, H, W, H, B, R)
    outs = []
    for lit in ("0", "1"):
        env = dict(os.environ, ME_ORACLE_LITERAL=lit)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, check=True,
                                   capture_output=True, text=True).stdout)
    assert outs[0] == outs[1] and len(outs[0]) > 0


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("B,R,W,H,seed", [
    (8, 12, 64, 48, 1), (16, 32, 96, 80, 2), (4, 15, 33, 29, 3), (5, 7, 41, 23, 4), (7, 1, 30, 30, 5),
    (16, 64, 48, 40, 6), (32, 15, 80, 72, 7), (64, 7, 130, 70, 8), (8, 0, 32, 32, 9), (3, 2, 7, 5, 10),
])
def test_oracle_vs_live_reference_random(orc, B, R, W, H, seed):
    """Randomised differential test against the unmodified reference functions."""
    from motionestimation_b200 import frames
    ref_lib = Ref()
    pairs = [frames.random_pair(W, H, seed), frames.shifted_noise_pair(W, H, seed=seed, shift=(2, -1)),
             frames.far_pair(W, H, seed)]
    for cur, ref in pairs:
        r = ref_lib.search(cur, ref, B, R)
        o = orc.search(cur, ref, B, R)
        assert np.array_equal(r["mvx"], o["mvx"]) and np.array_equal(r["mvy"], o["mvy"])
        assert np.array_equal(r["score"].view(np.uint32), o["score"].view(np.uint32))


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built")
def test_reference_pool_path_matches(orc):
    """The reference's own thread-pool dispatch (main.c:144-158) gives the same MVs."""
    from motionestimation_b200 import frames
    cur, ref = frames.foreman(2), frames.foreman(1)
    sec, r = Ref().search_pool(cur, ref, 8, 12, pool_threads=100)
    o = orc.search(cur, ref, 8, 12)
    assert sec > 0
    assert np.array_equal(r["mvx"], o["mvx"]) and np.array_equal(r["mvy"], o["mvy"])
    assert np.array_equal(r["score"], np.trunc(o["score"]))  # int val = float score, main.c:104


def test_work_counts(orc):
    """Exact pixel-compare / candidate counts."""
    from motionestimation_b200 import frames
    assert orc.num_blocks(352, 288, 8) == 1584
    assert orc.candidates(352, 288, 8, 12) == 927024
    assert orc.pixel_compares(352, 288, 8, 12) == 59329536
    assert orc.num_blocks(1920, 1080, 16) == 8160
    assert orc.candidates(1920, 1080, 16, 32) == 33188832
    assert orc.candidates(1920, 1080, 16, 64) == 127647200
    for (W, H, B, R) in [(1920, 1080, 16, 32), (3840, 2160, 8, 32), (37, 29, 5, 3)]:
        assert orc.pixel_compares(W, H, B, R) == frames.pixel_compares(W, H, B, R)
        assert orc.candidates(W, H, B, R) == frames.candidates(W, H, B, R)
    # brute-force count on a small case
    W, H, B, R = 37, 29, 5, 3
    x0, y0, w, h = frames.block_grid(W, H, B)
    pc = 0
    for i in range(len(x0)):
        ncx = min(W - 1, x0[i] + w[i] - 1 + R) - w[i] + 1 - max(0, x0[i] - R) + 1
        ncy = min(H - 1, y0[i] + h[i] - 1 + R) - h[i] + 1 - max(0, y0[i] - R) + 1
        pc += int(ncx * ncy * w[i] * h[i])
    assert pc == orc.pixel_compares(W, H, B, R)
