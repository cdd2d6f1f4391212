"""GPU tests of the three-step / diamond search kernels (run on the B200 box) through the C ABI.
PARITY UNPINNED against the reference (it has no fast search): the kernels are compared bit for
bit -- motion vectors, SSD, score bits, number of candidate evaluations -- with the definition in
oracle/me_oracle_fast.c, which tests/test_oracle_fast.py pins on the CPU."""
import os
import subprocess

import numpy as np
import pytest

import motionestimation_b200 as me
from oracle_binding import Oracle, TSS, DIAMOND, ROOT

pytestmark = pytest.mark.gpu

ALGOS = [(TSS, me.ME_SEARCH_THREE_STEP), (DIAMOND, me.ME_SEARCH_DIAMOND)]


@pytest.fixture(scope="module")
def orc():
    return Oracle()


def check(out, p, o, what):
    bad = np.nonzero((out["mvx"][p] != o["mvx"]) | (out["mvy"][p] != o["mvy"]) | (out["ssd"][p] != o["ssd"]))[0]
    assert bad.size == 0, f"{what}: {bad.size} mismatches at blocks {bad[:5]}: got " \
        f"({out['mvx'][p][bad[:5]]},{out['mvy'][p][bad[:5]]},{out['ssd'][p][bad[:5]]}) want {o[bad[:5]]}"
    assert np.array_equal(out["score"][p].view(np.uint32), o["score"].view(np.uint32)), f"{what}: score bits"


GEOMS = [
    (8, 7, 64, 48), (16, 7, 96, 80), (8, 12, 100, 60), (16, 32, 200, 104), (5, 3, 23, 17), (4, 15, 33, 29),
    (7, 9, 60, 41), (32, 16, 80, 72), (64, 8, 130, 70), (8, 0, 24, 16), (16, 15, 16, 16), (16, 64, 160, 144),
    (8, 1, 352, 16), (12, 5, 50, 37),
]


@pytest.mark.parametrize("algo,mode", ALGOS, ids=["three_step", "diamond"])
@pytest.mark.parametrize("B,R,W,H", GEOMS)
def test_fast_search_matches_definition(orc, B, R, W, H, algo, mode):
    pairs = [me.random_pair(W, H, B + R), me.shifted_noise_pair(W, H, seed=W + H, shift=(3, -2), cell=4),
             me.constant_pair(W, H), me.checker_pair(W, H, 2), me.far_pair(W, H, 1)]
    cur = np.stack([p[0] for p in pairs])
    ref = np.stack([p[1] for p in pairs])
    with me.Estimator(W, H, B, R, max_pairs=len(pairs), search=mode) as est:
        out = est.search_u8(cur, ref)
        evals = est.fast_evaluations
    want_evals = 0
    for p in range(len(pairs)):
        o, ev = orc.search_fast(cur[p], ref[p], B, R, algo)
        want_evals += ev
        check(out, p, o, f"pair {p}")
    assert evals == want_evals


@pytest.mark.parametrize("algo,mode", ALGOS, ids=["three_step", "diamond"])
def test_fast_search_foreman_and_drop_in(orc, algo, mode):
    """BASELINE config 4: Foreman, reference default block size / range, through the reference's
    own structs (me_b200_search_fast)."""
    cur8, ref8 = me.foreman(2), me.foreman(1)
    cur, ref = cur8.astype(np.int32).ravel(), ref8.astype(np.int32).ravel()
    pf = me.create_prediction_frame(cur, 352, 288, 8)
    sc, sd = me.search_prediction_frame(pf, ref, 12, want_scores=True, search=mode)
    o, _ = orc.search_fast(cur8, ref8, 8, 12, algo)
    assert np.array_equal(np.array([pf.blks[i].motion_vectorX for i in range(pf.num_blks)]), o["mvx"])
    assert np.array_equal(np.array([pf.blks[i].motion_vectorY for i in range(pf.num_blks)]), o["mvy"])
    assert all(pf.blks[i].is_best_match_found == 1 for i in range(pf.num_blks))
    assert np.array_equal(sd, o["ssd"]) and np.array_equal(sc.view(np.uint32), o["score"].view(np.uint32))


@pytest.mark.parametrize("algo,mode", ALGOS, ids=["three_step", "diamond"])
def test_fast_search_1080p_device_path(orc, algo, mode):
    """'Beauty' stand-in at 1080p, 16x16, +-32 on device-resident frames with a padded pitch,
    two block-row bands."""
    import torch
    W, H, B, R = 1920, 1080, 16, 32
    c, r = me.tiled_frames(W, H)
    pitch = 2048
    cur = torch.zeros((H, pitch), dtype=torch.uint8, device="cuda")
    ref = torch.zeros_like(cur)
    cur[:, :W] = torch.from_numpy(c).cuda()
    ref[:, :W] = torch.from_numpy(r).cuda()
    with me.Estimator(W, H, B, R, search=mode) as est:
        nb = est.num_blocks
        mvx = torch.full((nb,), -99, dtype=torch.int32, device="cuda")
        mvy = torch.full_like(mvx, -99)
        ssd = torch.zeros((nb,), dtype=torch.int32, device="cuda")
        score = torch.zeros((nb,), dtype=torch.float32, device="cuda")
        mid = est.blocks_y // 3
        est.search_device(cur, ref, pitch, 0, 1, mvx, mvy, ssd, score, by_begin=0, by_end=mid)
        est.search_device(cur, ref, pitch, 0, 1, mvx, mvy, ssd, score, by_begin=mid)
        torch.cuda.synchronize()
        evals = est.fast_evaluations
    o, ev = orc.search_fast(c, r, B, R, algo)
    out = {"mvx": mvx.cpu().numpy()[None], "mvy": mvy.cpu().numpy()[None],
           "ssd": ssd.cpu().numpy().astype(np.uint32)[None], "score": score.cpu().numpy()[None]}
    check(out, 0, o, "1080p")
    assert evals == ev


@pytest.mark.parametrize("algo,env", [(TSS, "tss"), (DIAMOND, "diamond")])
def test_fast_search_cli(tmp_path, orc, algo, env):
    """mes_b200 with ME_B200_SEARCH: the reference's argv, the fast pattern's field in mv_*.txt."""
    exe = os.path.join(ROOT, "motionestimation_b200", "mes_b200")
    g = os.path.join(ROOT, "motionestimation_b200", "data")   # the reference's frames/ directory
    p = subprocess.run([exe, f"{g}/ForemanYF2.yuv", f"{g}/ForemanYF1.yuv", str(tmp_path)],
                       capture_output=True, text=True, env=dict(os.environ, ME_B200_SEARCH=env))
    assert p.returncode == 0, p.stderr
    o, _ = orc.search_fast(me.foreman(2), me.foreman(1), 8, 12, algo)
    rows = [l.split() for l in open(tmp_path / "mv_8_12.txt")]
    assert [int(r[5]) for r in rows] == o["mvx"].tolist() and [int(r[6]) for r in rows] == o["mvy"].tolist()
    assert [int(r[7]) for r in rows] == o["ssd"].tolist()
    out5, psnr = orc.output5(me.foreman(2), me.foreman(1), 8, o)
    assert open(tmp_path / "output_8_12.yuv", "rb").read() == out5.tobytes()
    assert "PSNR: %.6f\n" % psnr in p.stdout
