"""CPU tests of the fast-search definition (oracle/me_oracle_fast.c).  PARITY UNPINNED: the
reference has no three-step / diamond search, so these tests pin the DEFINITION instead:
an independent pure-Python restatement of the same rules, and the properties any such
search must have relative to the reference's exhaustive search."""
import numpy as np
import pytest

from motionestimation_b200 import frames
from oracle_binding import Oracle, TSS, DIAMOND

SQUARE8 = [(-1, -1), (0, -1), (1, -1), (-1, 0), (1, 0), (-1, 1), (0, 1), (1, 1)]
LARGE8 = [(0, -2), (-1, -1), (1, -1), (-2, 0), (2, 0), (-1, 1), (1, 1), (0, 2)]
SMALL4 = [(0, -1), (-1, 0), (1, 0), (0, 1)]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def py_fast(cur, ref, B, R, algo, first_step):
    """Independent restatement: incumbent centre, pattern points in raster order, strict '<'."""
    H, W = cur.shape
    cur = cur.astype(np.int64)
    ref = ref.astype(np.int64)
    x0s, y0s, ws, hs = frames.block_grid(W, H, B)
    res, evals = [], 0
    for x0, y0, w, h in zip(x0s, y0s, ws, hs):
        lo_x, lo_y = max(0, x0 - R), max(0, y0 - R)
        hi_x = min(W - 1, x0 + w - 1 + R) - w + 1
        hi_y = min(H - 1, y0 + h - 1 + R) - h + 1
        blk = cur[y0:y0 + h, x0:x0 + w]

        def score(x, y):
            ssd = int(((blk - ref[y:y + h, x:x + w]) ** 2).sum())
            assert ssd < 1 << 24
            return np.float32(ssd) / np.float32(w * h), ssd

        bx, by = x0, y0
        best, bssd = score(bx, by)
        evals += 1

        def visit(offs, scale):
            nonlocal bx, by, best, bssd, evals
            cx, cy = bx, by
            for ox, oy in offs:
                x, y = cx + ox * scale, cy + oy * scale
                if lo_x <= x <= hi_x and lo_y <= y <= hi_y:
                    s, d = score(x, y)
                    evals += 1
                    if s < best:
                        best, bssd, bx, by = s, d, x, y

        if algo == TSS:
            st = first_step
            while st >= 1:
                visit(SQUARE8, st)
                st //= 2
        else:
            while True:
                cx, cy = bx, by
                visit(LARGE8, 1)
                if (bx, by) == (cx, cy):
                    break
            visit(SMALL4, 1)
        res.append((bx - x0, by - y0, bssd, best))
    return res, evals


@pytest.mark.parametrize("algo", [TSS, DIAMOND], ids=["three_step", "diamond"])
@pytest.mark.parametrize("B,R,W,H,seed", [(8, 7, 64, 48, 1), (16, 7, 96, 80, 2), (8, 12, 50, 37, 3), (5, 3, 23, 17, 4),
                                          (16, 32, 80, 64, 5), (4, 15, 33, 29, 6), (8, 0, 24, 16, 7)])
def test_fast_oracle_matches_python_restatement(orc, algo, B, R, W, H, seed):
    for cur, ref in (frames.shifted_noise_pair(W, H, seed=seed, shift=(2, -1), cell=4),
                     frames.random_pair(W, H, seed), frames.constant_pair(W, H)):
        got, ev = orc.search_fast(cur, ref, B, R, algo)
        want, ev_want = py_fast(cur, ref, B, R, algo, orc.lib.me_oracle_tss_first_step(R))
        assert ev == ev_want
        assert [tuple(int(v) for v in (g["mvx"], g["mvy"], g["ssd"])) for g in got] == [w[:3] for w in want]
        assert np.array_equal(got["score"], np.array([w[3] for w in want], np.float32))


@pytest.mark.parametrize("algo", [TSS, DIAMOND], ids=["three_step", "diamond"])
def test_fast_search_properties(orc, algo):
    """On Foreman: every MV stays inside the clamped window, the cost is never better than the
    reference's exhaustive minimum and never worse than zero motion, and it equals the cost of
    the reported MV under the reference's cost function."""
    cur, ref = frames.foreman(2), frames.foreman(1)
    B, R = 8, 7
    fast, ev = orc.search_fast(cur, ref, B, R, algo)
    full = orc.search(cur, ref, B, R)
    x0, y0, w, h = frames.block_grid(352, 288, B)
    assert np.all(fast["ssd"] >= full["ssd"])
    assert np.all(x0 + fast["mvx"] >= np.maximum(0, x0 - R)) and np.all(x0 + fast["mvx"] + w <= 352)
    assert np.all(np.abs(fast["mvx"]) <= R) and np.all(np.abs(fast["mvy"]) <= R)
    c, r = cur.astype(np.int64), ref.astype(np.int64)
    for i in range(0, len(fast), 37):
        yy, xx = y0[i] + fast["mvy"][i], x0[i] + fast["mvx"][i]
        ssd = int(((c[y0[i]:y0[i] + h[i], x0[i]:x0[i] + w[i]] - r[yy:yy + h[i], xx:xx + w[i]]) ** 2).sum())
        zero = int(((c[y0[i]:y0[i] + h[i], x0[i]:x0[i] + w[i]] - r[y0[i]:y0[i] + h[i], x0[i]:x0[i] + w[i]]) ** 2).sum())
        assert ssd == fast["ssd"][i] <= zero
    # a fast search looks at a small fraction of the (2R+1)^2 candidates
    assert ev < 0.2 * orc.candidates(352, 288, B, R)
    # and still finds most of the exhaustive minima on real video
    assert np.mean(fast["ssd"] == full["ssd"]) > 0.5


def test_three_step_first_step(orc):
    f = orc.lib.me_oracle_tss_first_step
    assert [f(r) for r in (0, 1, 2, 3, 7, 8, 12, 15, 16, 32, 64)] == [0, 1, 1, 2, 4, 4, 4, 8, 8, 16, 32]
