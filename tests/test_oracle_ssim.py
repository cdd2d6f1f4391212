"""CPU tests of the SSIM checker (no GPU): oracle/me_oracle_ssim.c is pinned against
(1) fixtures produced by the unmodified reference SSIM search (tests/golden/golden_ssim.json,
    fields_ssim.npz, made by tests/golden/make_golden_ssim.py),
(2) the unmodified reference run live when oracle/_ref/libme_ref_ssim.so is present,
(3) its own literal float-accumulation mode, and
(4) the one number the reference's logs hold for this code: "Original Score: 384.4514"
    (results/cpu/foreman/4_15.txt:10, 4_7.txt:10), the float accumulation of main_ssim.c:88-95."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from cases import SSIM_CASES, make_frames, load_golden_ssim
from oracle_binding import Oracle, RefSsim, ROOT

META, FIELDS = load_golden_ssim()


@pytest.fixture(scope="module")
def orc():
    return Oracle()


@pytest.mark.parametrize("case", SSIM_CASES, ids=[c[0] for c in SSIM_CASES])
def test_ssim_oracle_matches_reference_fixture(orc, case):
    name, gen, args, B, R = case
    cur, ref = make_frames(gen, args)
    m = META[name]
    assert hashlib.md5(cur.tobytes()).hexdigest() == m["cur_md5"], "synthetic generator drifted"
    assert hashlib.md5(ref.tobytes()).hexdigest() == m["ref_md5"]
    o = orc.search_ssim(cur, ref, B, R)
    assert len(o) == m["blocks"]
    assert np.array_equal(o["mvx"], FIELDS[name + "/mvx"])
    assert np.array_equal(o["mvy"], FIELDS[name + "/mvy"])
    assert np.array_equal(o["score"].view(np.uint32), FIELDS[name + "/score_bits"])
    assert np.array_equal(o["ssd"], FIELDS[name + "/found"])
    assert int(np.count_nonzero(o["ssd"] == 0)) == m["not_found"]
    # post-processing of main_ssim.c:80-95 on that field: planes + the printed scores
    out5, _ = orc.output5(cur, ref, B, o)
    assert hashlib.md5(out5.tobytes()).hexdigest() == m["yuv_md5"]
    H = cur.shape[0]
    orig, comp = orc.ssim_frame_scores(cur, ref, out5[2 * H:3 * H])
    assert "Original Score: %.4f, Compensated Score: %.4f" % (orig, comp) == m["scores_line"]


def test_logged_original_score(orc):
    """results/cpu/foreman/4_15.txt:10 and 4_7.txt:10 log 'Original Score: 384.4514' for
    cur = ForemanYF4, ref = ForemanYF1; with the MSE field of B=4 R=15 / R=7 the same logs give
    'Compensated Score: 17.1327' / '47.7937'."""
    from motionestimation_b200 import frames
    cur, ref = frames.foreman(4), frames.foreman(1)
    for R, comp_want in ((15, "17.1327"), (7, "47.7937")):
        o = orc.search(cur, ref, 4, R)
        out5, _ = orc.output5(cur, ref, 4, o)
        orig, comp = orc.ssim_frame_scores(cur, ref, out5[2 * 288:3 * 288])
        assert "%.4f" % orig == "384.4514"
        assert "%.4f" % comp == comp_want
    assert META["ssim_foreman_yf4_yf1_4_15"]["scores_line"].startswith("Original Score: 384.4514,")


def test_no_positive_candidate_is_reported(orc):
    """Anti-correlated frames: the reference never writes the MV (ssim.c:88-103); the checker
    (like the harness, which zero-fills malloc) reports (0,0), score 0, found 0."""
    from motionestimation_b200 import frames
    cur, ref = frames.inverted_pair(96, 80, seed=21)
    o = orc.search_ssim(cur, ref, 16, 3)
    assert not o["ssd"].any() and not o["score"].any() and not o["mvx"].any() and not o["mvy"].any()


def test_constant_frames_first_candidate_wins(orc):
    """Zero variance everywhere: every candidate scores exactly 1.0, strict '>' keeps the first
    one in y-major/x-minor order (ssim.c:98-106) => MV = (-min(R,x0), -min(R,y0))."""
    from motionestimation_b200 import frames
    cur, ref = frames.constant_pair(96, 64)
    o = orc.search_ssim(cur, ref, 8, 12)
    x0, y0, _, _ = frames.block_grid(96, 64, 8)
    assert np.array_equal(o["mvx"], -np.minimum(12, x0))
    assert np.array_equal(o["mvy"], -np.minimum(12, y0))
    assert np.all(o["score"] == 1.0) and np.all(o["ssd"] == 1)


@pytest.mark.parametrize("B,R,W,H", [(8, 3, 40, 24), (16, 5, 50, 37), (24, 4, 60, 50), (5, 2, 23, 17)])
def test_ssim_literal_accumulation_agrees(B, R, W, H):
    """The integer pixel-sum / cross-sum shortcuts equal the literal float loops of ssim.c."""
    code = (
        "import sys, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from oracle_binding import Oracle; from motionestimation_b200 import frames\n"
        "o = Oracle(); out = []\n"
        "for cur, ref in (frames.far_pair(%d, %d, 3), frames.random_pair(%d, %d, 4),\n"
        "                 frames.shifted_noise_pair(%d, %d, seed=5, shift=(1, 2))):\n"
        "    r = o.search_ssim(cur, ref, %d, %d, nthreads=4)\n"
        "    out.append(np.concatenate([r['mvx'], r['mvy'], r['score'].view(np.int32)]))\n"
        "sys.stdout.write(' '.join(map(str, np.concatenate(out).tolist())))\n"
    ) % (ROOT, os.path.join(ROOT, "tests"), W, H, W, H, W, H, B, R)
    outs = []
    for lit in ("0", "1"):
        env = dict(os.environ, ME_ORACLE_LITERAL=lit)
        outs.append(subprocess.run([sys.executable, "-c", code], env=env, check=True,
                                   capture_output=True, text=True).stdout)
    assert outs[0] == outs[1] and len(outs[0]) > 0


@pytest.mark.skipif(not RefSsim.available(), reason="oracle/_ref/libme_ref_ssim.so not built (needs /root/reference)")
@pytest.mark.parametrize("B,R,W,H,seed", [
    (8, 12, 64, 48, 1), (16, 7, 96, 80, 2), (4, 15, 33, 29, 3), (5, 7, 41, 23, 4), (7, 1, 30, 30, 5),
    (16, 32, 48, 40, 6), (32, 9, 80, 72, 7), (64, 5, 130, 70, 8), (8, 0, 32, 32, 9), (3, 2, 7, 5, 10),
])
def test_ssim_oracle_vs_live_reference_random(orc, B, R, W, H, seed):
    """Randomised differential test against the unmodified findBestBlkSSIM."""
    from motionestimation_b200 import frames
    ref_lib = RefSsim()
    pairs = [frames.random_pair(W, H, seed), frames.shifted_noise_pair(W, H, seed=seed, shift=(2, -1)),
             frames.far_pair(W, H, seed), frames.inverted_pair(W, H, seed=seed, period=11.0)]
    for cur, ref in pairs:
        r = ref_lib.search(cur, ref, B, R)
        o = orc.search_ssim(cur, ref, B, R)
        for k in ("mvx", "mvy", "ssd"):
            assert np.array_equal(r[k], o[k]), k
        assert np.array_equal(r["score"].view(np.uint32), o["score"].view(np.uint32))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "mes_ref_ssim")),
                    reason="reference SSIM binary not built")
def test_reference_ssim_binary_matches_fixture(tmp_path):
    """The stand-alone reference program, built exactly like src/cpu/run_ssim.sh:4 (no -O),
    writes the yuv and prints the scores line the fixtures hold (built -O2 in the harness)."""
    g = os.path.join(ROOT, "motionestimation_b200", "data")   # the reference's frames/ directory
    out = subprocess.run([os.path.join(ROOT, "oracle", "_ref", "mes_ref_ssim"), os.path.join(g, "ForemanYF4.yuv"),
                          os.path.join(g, "ForemanYF1.yuv"), str(tmp_path), "16", "7", "352", "288"],
                         capture_output=True, text=True, check=True).stdout
    m = META["ssim_foreman_yf4_yf1_16_7"]
    assert m["scores_line"] in out
    assert hashlib.md5(open(tmp_path / "output_16_7.yuv", "rb").read()).hexdigest() == m["yuv_md5"]
