"""CPU tests of the plain-C host layer (include/me_common.h) against the oracle
restatement and, when built, the unmodified reference: block grid, .yuv I/O,
motion compensation, frame difference, PSNR."""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np
import pytest

import motionestimation_b200 as me
from cases import load_golden
from oracle_binding import Oracle, Ref

META, FIELDS = load_golden()
IP = C.POINTER(C.c_int)


def iptr(a):
    return a.ctypes.data_as(IP)


@pytest.mark.parametrize("W,H,B", [(352, 288, 8), (1920, 1080, 16), (37, 29, 5), (16, 16, 16), (7, 5, 3), (10, 10, 64)])
def test_block_grid_matches_oracle(W, H, B):
    """createPredictionFrame: raster order, partial edge blocks (prediction_frame.c:9-23)."""
    cur = np.zeros(W * H, np.int32)
    pf = me.create_prediction_frame(cur, W, H, B)
    x0, y0, w, h = me.block_grid(W, H, B)
    assert pf.num_blks == len(x0) == Oracle().num_blocks(W, H, B)
    assert (pf.width, pf.height, pf.blk_dim) == (W, H, B)
    nbx = -(-W // B)
    for i in range(pf.num_blks):
        b = pf.blks[i]
        assert (b.idx_x, b.idx_y) == (i % nbx, i // nbx)
        assert (b.top_left_x, b.top_left_y, b.width, b.height) == (x0[i], y0[i], w[i], h[i])
        assert (b.bottom_right_x, b.bottom_right_y) == (x0[i] + w[i] - 1, y0[i] + h[i] - 1)
        assert b.motion_vectorY == -1000 and b.is_best_match_found == 0  # block.c:12


def test_yuv_read_write_roundtrip(tmp_path):
    lib = me.load_library()
    g = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "motionestimation_b200", "data", "ForemanYF1.yuv")
    n = 352 * 288
    buf = np.zeros(n, np.int32)
    assert lib.yuvReadFrame(g.encode(), iptr(buf), n) == 1
    raw = np.frombuffer(open(g, "rb").read(), np.uint8)
    assert np.array_equal(buf, raw.astype(np.int32))           # utils.c:49-53 widening
    u8 = np.zeros(n, np.uint8)
    assert lib.yuvReadFrameU8(g.encode(), u8.ctypes.data_as(C.c_void_p), n) == 1
    assert np.array_equal(u8, raw)
    out = tmp_path / "o.yuv"
    buf2 = buf.copy()
    buf2[:4] = [256, 257, -1, 511]                               # C cast narrowing, utils.c:55-59
    assert lib.yuvWriteFrame(str(out).encode(), iptr(buf2), n) == 1
    back = np.frombuffer(open(out, "rb").read(), np.uint8)
    assert list(back[:4]) == [0, 1, 255, 255] and np.array_equal(back[4:], raw[4:])
    # failure conventions: 0, no crash (the reference dereferences NULL here, utils.c:62-64)
    assert lib.yuvReadFrame(b"/nonexistent/file.yuv", iptr(buf), n) == 0
    assert lib.yuvReadFrame(g.encode(), iptr(np.zeros(n + 10, np.int32)), n + 10) == 0  # short file
    assert lib.yuvWriteFrame(b"/nonexistent/dir/o.yuv", iptr(buf), n) == 0


@pytest.mark.parametrize("name", ["foreman_yf4_yf1_8_12", "foreman_yf4_yf1_4_15", "foreman_yf2_yf1_5_7",
                                  "noise_200x120_16_32", "foreman_yf2_yf1_64_8"])
def test_postprocessing_reproduces_golden_yuv(name):
    """motionCompensatedFrame + frameDiff + imagePSNR on the golden MV field give
    the reference's output_<B>_<R>.yuv bytes and PSNR line (main.c:160-171)."""
    from cases import CASES, make_frames
    lib = me.load_library()
    case = [c for c in CASES if c[0] == name][0]
    cur8, ref8 = make_frames(case[1], case[2])
    B = case[3]
    H, W = cur8.shape
    n = W * H
    cur, ref = cur8.astype(np.int32).ravel(), ref8.astype(np.int32).ravel()
    pf = me.create_prediction_frame(cur, W, H, B)
    out = np.zeros(5 * n, np.int32)
    out[:n], out[n:2 * n] = ref, cur
    # no vectors yet -> refused (the reference exit(0)s, utils.c:105-108)
    assert lib.motionCompensatedFrame(iptr(out[2 * n:]), pf, iptr(ref)) == 0
    for i in range(pf.num_blks):
        pf.blks[i].motion_vectorX = int(FIELDS[name + "/mvx"][i])
        pf.blks[i].motion_vectorY = int(FIELDS[name + "/mvy"][i])
        pf.blks[i].is_best_match_found = 1
    mc = out[2 * n:3 * n]
    assert lib.motionCompensatedFrame(iptr(mc), pf, iptr(ref)) == 1
    d0, d1 = out[3 * n:4 * n], out[4 * n:]
    lib.frameDiff(iptr(d0), iptr(ref), iptr(cur), n)
    lib.frameDiff(iptr(d1), iptr(mc), iptr(cur), n)
    psnr = lib.imagePSNR(iptr(mc), iptr(cur), W, H)
    assert "%.6f" % psnr == META[name]["psnr"]
    assert hashlib.md5(out.astype(np.uint8).tobytes()).hexdigest() == META[name]["yuv_md5"]
    if Ref.available():
        ro, rp = Ref().output5(cur8, ref8, B, FIELDS[name + "/mvx"], FIELDS[name + "/mvy"])
        assert rp == psnr and np.array_equal(ro.ravel(), out.astype(np.uint8))


def test_psnr_identical_frames_is_99():
    lib = me.load_library()
    a = np.full(64, 7, np.int32)
    assert lib.imagePSNR(iptr(a), iptr(a.copy()), 8, 8) == 99.0   # utils.c:160


def test_timestamp_monotone():
    lib = me.load_library()
    t0 = lib.getTimeStamp()
    t1 = lib.getTimeStamp()
    assert t1 >= t0 > 1.0e9


def test_int_to_u8_narrowing_matches_numpy(tmp_path):
    """host/me_pack.c: the narrowing of the reference's int pixels (utils.c:49-53 in reverse) at the
    drop-in seam -- SSE2/AVX2 body + scalar tail, every length phase; returns the OR of the inputs so
    the caller can reject pixels outside 0..255."""
    # an internal function of libme_b200.so (not part of the ABI): compiled on its own for the test
    src_c = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "motionestimation_b200",
                         "host", "me_pack.c")
    so = str(tmp_path / "me_pack.so")
    subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src_c], check=True)
    fn = C.CDLL(so).me_pack_int_to_u8
    fn.restype = C.c_uint
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.Generator(np.random.PCG64(5))
    for n in list(range(0, 70)) + [352 * 288, 1920 * 135 + 7]:
        src = rng.integers(0, 256, n, dtype=np.int32)
        dst = np.full(n + 8, 0xAB, np.uint8)
        acc = fn(dst.ctypes.data, src.ctypes.data, n)
        assert np.array_equal(dst[:n], src.astype(np.uint8)) and np.all(dst[n:] == 0xAB)
        assert acc == (int(np.bitwise_or.reduce(src)) if n else 0)
    for bad in (256, -1, 70000, -40000):
        src = rng.integers(0, 256, 100, dtype=np.int32)
        src[57] = bad
        dst = np.zeros(100, np.uint8)
        assert fn(dst.ctypes.data, src.ctypes.data, 100) & ~0xff
