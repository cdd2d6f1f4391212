"""GPU parity tests of the SSIM-cost search (run on the B200 box): every call goes through the
C ABI of libme_b200.so (context mode ME_COST_SSIM, or the me_b200_search_ssim drop-in) and is
compared bit for bit -- motion vectors, found flag, float score bits -- with the fixtures made
by the unmodified reference (tests/golden/*_ssim.*) and with the pinned restatement."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

import motionestimation_b200 as me
from cases import SSIM_CASES, make_frames, load_golden_ssim
from oracle_binding import Oracle, ROOT

pytestmark = pytest.mark.gpu

META, FIELDS = load_golden_ssim()
KERNELS = [me.ME_KERNEL_GENERIC, me.ME_KERNEL_AUTO]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def check(out, p, exp_mvx, exp_mvy, exp_found, exp_bits, what=""):
    bad = np.nonzero((out["mvx"][p] != exp_mvx) | (out["mvy"][p] != exp_mvy) |
                     (out["score"][p].view(np.uint32) != exp_bits))[0]
    assert bad.size == 0, f"{what}: {bad.size} mismatches, first blocks {bad[:5]}: got " \
        f"({out['mvx'][p][bad[:5]]},{out['mvy'][p][bad[:5]]},{out['score'][p][bad[:5]]}) want " \
        f"({exp_mvx[bad[:5]]},{exp_mvy[bad[:5]]},{exp_bits[bad[:5]].view(np.float32)})"
    assert np.array_equal(out["ssd"][p], exp_found), f"{what}: found flags differ"


@pytest.mark.parametrize("kernel", KERNELS, ids=["generic", "auto"])
@pytest.mark.parametrize("case", SSIM_CASES, ids=[c[0] for c in SSIM_CASES])
def test_ssim_golden_cases(case, kernel):
    name, gen, args, B, R = case
    cur, ref = make_frames(gen, args)
    H, W = cur.shape
    with me.Estimator(W, H, B, R, kernel=kernel, cost=me.ME_COST_SSIM) as est:
        assert est.num_blocks == META[name]["blocks"]
        out = est.search_u8(cur, ref)
        assert est.launch_count >= 1
    check(out, 0, FIELDS[name + "/mvx"].astype(np.int32), FIELDS[name + "/mvy"].astype(np.int32),
          FIELDS[name + "/found"].astype(np.uint32), FIELDS[name + "/score_bits"], name)


GEOMS = [
    # B, R, W, H
    (8, 12, 64, 48), (8, 12, 100, 60), (16, 7, 96, 80), (16, 32, 200, 104), (16, 64, 160, 144),
    (8, 32, 128, 72), (4, 15, 33, 29), (5, 7, 41, 23), (7, 1, 30, 30), (32, 9, 80, 72),
    (8, 0, 32, 32), (3, 2, 7, 5), (16, 32, 16, 16), (8, 12, 8, 8), (16, 8, 48, 40),
    (16, 7, 100, 50), (8, 3, 70, 45), (16, 0, 64, 64), (8, 1, 16, 8), (16, 3, 300, 70), (8, 13, 88, 60),
    (16, 33, 128, 104), (16, 120, 64, 48), (8, 5, 352, 16),
]


@pytest.mark.parametrize("kernel", KERNELS, ids=["generic", "auto"])
@pytest.mark.parametrize("B,R,W,H", GEOMS)
def test_ssim_random_differential(orc, B, R, W, H, kernel):
    pairs = [me.random_pair(W, H, B + R), me.shifted_noise_pair(W, H, seed=W + H, shift=(3, -2)),
             me.constant_pair(W, H), me.checker_pair(W, H, 2), me.far_pair(W, H, 1),
             me.inverted_pair(W, H, seed=R, period=9.0)]
    cur = np.stack([p[0] for p in pairs])
    ref = np.stack([p[1] for p in pairs])
    with me.Estimator(W, H, B, R, max_pairs=len(pairs), kernel=kernel, cost=me.ME_COST_SSIM) as est:
        out = est.search_u8(cur, ref)
    for p in range(len(pairs)):
        o = orc.search_ssim(cur[p], ref[p], B, R)
        check(out, p, o["mvx"], o["mvy"], o["ssd"], o["score"].view(np.uint32), f"pair {p}")


F4_GEOMS = [g for g in GEOMS if g[0] == 16] + [(16, 32, 400, 200), (16, 12, 330, 90), (16, 5, 1000, 64),
                                               (16, 40, 256, 256), (16, 1, 64, 48), (16, 17, 48, 300), (16, 9, 349, 55)]


@pytest.mark.parametrize("B,R,W,H", F4_GEOMS)
def test_ssim_on_the_tiled_kernel(orc, monkeypatch, B, R, W, H):
    """ME_B200_SSIM_FORM4=1: the SSIM cost as a formulation of the MSE path's tiled kernel (me_tiled.cu FORM 4; opt-in,
    measured slower than the streaming kernel -- DESIGN.md 5.5).  Partial-width right column, odd spans (table tile
    phase), odd widths (falls back), windows that do not fit (falls back), constant frames (every candidate ties)."""
    monkeypatch.setenv("ME_B200_SSIM_FORM4", "1")
    pairs = [me.random_pair(W, H, B + R), me.shifted_noise_pair(W, H, seed=W + H, shift=(3, -2)),
             me.constant_pair(W, H), me.checker_pair(W, H, 2), me.inverted_pair(W, H, seed=R, period=9.0),
             me.far_pair(W, H, 3)]
    cur = np.stack([p[0] for p in pairs])
    ref = np.stack([p[1] for p in pairs])
    with me.Estimator(W, H, B, R, max_pairs=len(pairs), cost=me.ME_COST_SSIM) as est:
        out = est.search_u8(cur, ref)
    for p in range(len(pairs)):
        o = orc.search_ssim(cur[p], ref[p], B, R)
        check(out, p, o["mvx"], o["mvy"], o["ssd"], o["score"].view(np.uint32), f"pair {p}")


def test_ssim_drop_in_prediction_frame():
    """me_b200_search_ssim on the reference's own structs (replaces main_ssim.c:67-77)."""
    cur8, ref8 = me.foreman(4), me.foreman(1)
    cur, ref = cur8.astype(np.int32).ravel(), ref8.astype(np.int32).ravel()
    pf = me.create_prediction_frame(cur, 352, 288, 16)
    sc, found = me.search_prediction_frame(pf, ref, 7, want_scores=True, cost=me.ME_COST_SSIM)
    name = "ssim_foreman_yf4_yf1_16_7"
    mvx = np.array([pf.blks[i].motion_vectorX for i in range(pf.num_blks)])
    mvy = np.array([pf.blks[i].motion_vectorY for i in range(pf.num_blks)])
    assert all(pf.blks[i].is_best_match_found == 1 for i in range(pf.num_blks))
    assert np.array_equal(mvx, FIELDS[name + "/mvx"]) and np.array_equal(mvy, FIELDS[name + "/mvy"])
    assert np.array_equal(sc.view(np.uint32), FIELDS[name + "/score_bits"])
    assert np.array_equal(found, FIELDS[name + "/found"])
    # the MSE drop-in on the same structs still answers with the MSE field (separate cached context)
    me.search_prediction_frame(pf, ref, 7)
    mse = Oracle().search(cur8, ref8, 16, 7)
    assert np.array_equal(np.array([pf.blks[i].motion_vectorX for i in range(pf.num_blks)]), mse["mvx"])


def test_ssim_device_path_bands_and_batches(orc):
    """Device-resident entry point: a batch in one call, two block-row bands, padded pitch."""
    import torch
    W, H, B, R = 208, 120, 16, 12
    pairs = [me.shifted_noise_pair(W, H, seed=s, shift=(s, -s)) for s in (1, 2, 3)]
    exp = [orc.search_ssim(c, r, B, R) for c, r in pairs]
    pitch = 256
    cur = torch.zeros((3, H, pitch), dtype=torch.uint8, device="cuda")
    ref = torch.zeros_like(cur)
    for i, (c, r) in enumerate(pairs):
        cur[i, :, :W] = torch.from_numpy(c).cuda()
        ref[i, :, :W] = torch.from_numpy(r).cuda()
    with me.Estimator(W, H, B, R, max_pairs=3, cost=me.ME_COST_SSIM) as est:
        nb = est.num_blocks
        mvx = torch.full((3, nb), -99, dtype=torch.int32, device="cuda")
        mvy = torch.full_like(mvx, -99)
        found = torch.zeros((3, nb), dtype=torch.int32, device="cuda")
        score = torch.zeros((3, nb), dtype=torch.float32, device="cuda")
        mid = est.blocks_y // 2
        est.search_device(cur, ref, pitch, pitch * H, 3, mvx, mvy, found, score, by_begin=0, by_end=mid)
        est.search_device(cur, ref, pitch, pitch * H, 3, mvx, mvy, found, score, by_begin=mid)
        torch.cuda.synchronize()
    out = {"mvx": mvx.cpu().numpy(), "mvy": mvy.cpu().numpy(), "ssd": found.cpu().numpy().astype(np.uint32),
           "score": score.cpu().numpy()}
    for p in range(3):
        check(out, p, exp[p]["mvx"], exp[p]["mvy"], exp[p]["ssd"], exp[p]["score"].view(np.uint32), f"pair {p}")


@pytest.mark.parametrize("W,H,B,R", [(1920, 1080, 16, 7), (1920, 1080, 16, 32), (1920, 1080, 8, 12),
                                     (3840, 2160, 16, 7)])   # the last one = main_ssim.c's own defaults
def test_ssim_full_size(orc, monkeypatch, W, H, B, R):
    """Full-size frames (1080p: half-height bottom row at B = 16; 4K at the SSIM program's default
    block size and span): the tuned path equals the pinned restatement on the first, a middle and
    the last block rows, and equals the generic kernel everywhere."""
    cur, ref = me.tiled_frames(W, H)
    with me.Estimator(W, H, B, R, cost=me.ME_COST_SSIM) as est:
        out = est.search_u8(cur, ref)
        nbx, nby = est.blocks_x, est.blocks_y
    if B == 16:   # and the opt-in formulation on the tiled kernel
        monkeypatch.setenv("ME_B200_SSIM_FORM4", "1")
        with me.Estimator(W, H, B, R, cost=me.ME_COST_SSIM) as est:
            out4 = est.search_u8(cur, ref)
        monkeypatch.delenv("ME_B200_SSIM_FORM4")
        for k in ("mvx", "mvy", "ssd"):
            assert np.array_equal(out[k], out4[k]), k
        assert np.array_equal(out["score"].view(np.uint32), out4["score"].view(np.uint32))
    with me.Estimator(W, H, B, R, kernel=me.ME_KERNEL_GENERIC, cost=me.ME_COST_SSIM) as est:
        gen = est.search_u8(cur, ref)
    for k in ("mvx", "mvy", "ssd"):
        assert np.array_equal(out[k], gen[k]), k
    assert np.array_equal(out["score"].view(np.uint32), gen["score"].view(np.uint32))
    for by in (0, nby // 2, nby - 1):
        o = orc.search_ssim(cur, ref, B, R, begin=by * nbx, end=(by + 1) * nbx)
        sl = slice(by * nbx, (by + 1) * nbx)
        sub = {k: v[:, sl] for k, v in out.items()}
        check(sub, 0, o["mvx"], o["mvy"], o["ssd"], o["score"].view(np.uint32), f"block row {by}")


def test_ssim_mode_rules():
    with me.Estimator(64, 48, 8, 4) as est:
        lib = me.load_library()
        assert lib.me_b200_set_cost(est._h, 7) == me.ME_ERR_INVALID_ARG
        assert lib.me_b200_set_search(est._h, me.ME_SEARCH_DIAMOND) == me.ME_OK
        assert lib.me_b200_set_cost(est._h, me.ME_COST_SSIM) == me.ME_ERR_UNSUPPORTED  # SSIM x fast pattern
        assert lib.me_b200_set_search(est._h, me.ME_SEARCH_FULL) == me.ME_OK
        assert lib.me_b200_set_cost(est._h, me.ME_COST_SSIM) == me.ME_OK
        assert lib.me_b200_set_search(est._h, me.ME_SEARCH_THREE_STEP) == me.ME_ERR_UNSUPPORTED


@pytest.mark.parametrize("args,name", [(("16", "7", "352", "288"), "ssim_foreman_yf4_yf1_16_7"),
                                       (("4", "15", "352", "288"), "ssim_foreman_yf4_yf1_4_15")])
def test_ssim_cli_is_byte_identical(tmp_path, args, name):
    """mes_b200_ssim: same argv, same 'Original Score ... Compensated Score' line and the same
    output_<B>_<R>.yuv bytes as the unmodified reference program (fixtures made from it)."""
    exe = os.path.join(ROOT, "motionestimation_b200", "mes_b200_ssim")
    g = os.path.join(ROOT, "motionestimation_b200", "data")   # the reference's frames/ directory
    p = subprocess.run([exe, f"{g}/ForemanYF4.yuv", f"{g}/ForemanYF1.yuv", str(tmp_path), *args],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    m = META[name]
    assert m["scores_line"] + "\n" in p.stdout
    assert "Output file dimensions: (352 x 1440)" in p.stdout
    out = open(tmp_path / f"output_{m['B']}_{m['R']}.yuv", "rb").read()
    assert hashlib.md5(out).hexdigest() == m["yuv_md5"]
    rows = [l.split() for l in open(tmp_path / f"mv_ssim_{m['B']}_{m['R']}.txt")]
    assert [int(r[5]) for r in rows] == FIELDS[name + "/mvx"].tolist()
    assert [int(r[6]) for r in rows] == FIELDS[name + "/mvy"].tolist()
    assert [int(r[8], 16) for r in rows] == FIELDS[name + "/score_bits"].tolist()
