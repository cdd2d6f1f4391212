"""ctypes access to the CPU checker under oracle/ -- TEST INFRASTRUCTURE ONLY.
`Oracle` = the C restatement (oracle/me_oracle.c); `Ref` = the unmodified
reference compiled into oracle/_ref/libme_ref.so (oracle/ref_harness.c), present
only where it was built (the container with /root/reference, or shipped prebuilt)."""
import ctypes as C
import hashlib
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libme_oracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libme_ref.so")
REF_SSIM_SO = os.path.join(ROOT, "oracle", "_ref", "libme_ref_ssim.so")
TSS, DIAMOND = 1, 2

RESULT_DTYPE = np.dtype([("mvx", np.int32), ("mvy", np.int32), ("ssd", np.uint32), ("score", np.float32)])


def _u8(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a, a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(ORACLE_SO)
        L = self.lib
        L.me_oracle_num_blocks.restype = C.c_int
        L.me_oracle_num_blocks.argtypes = [C.c_int] * 3
        L.me_oracle_search.restype = C.c_int
        L.me_oracle_search.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.me_oracle_pixel_compares.restype = C.c_uint64
        L.me_oracle_pixel_compares.argtypes = [C.c_int] * 4
        L.me_oracle_candidates.restype = C.c_uint64
        L.me_oracle_candidates.argtypes = [C.c_int] * 4
        L.me_oracle_motion_compensate.restype = C.c_int
        L.me_oracle_motion_compensate.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.me_oracle_frame_diff.restype = None
        L.me_oracle_frame_diff.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.me_oracle_psnr.restype = C.c_double
        L.me_oracle_psnr.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.me_oracle_search_ssim.restype = C.c_int
        L.me_oracle_search_ssim.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.me_oracle_ssim_frame_scores.restype = None
        L.me_oracle_ssim_frame_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                                  C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.me_oracle_search_fast.restype = C.c_int
        L.me_oracle_search_fast.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 8 + [C.c_void_p,
                                                                                       C.POINTER(C.c_uint64)]
        L.me_oracle_tss_first_step.restype = C.c_int
        L.me_oracle_tss_first_step.argtypes = [C.c_int]

    def num_blocks(self, W, H, B):
        return self.lib.me_oracle_num_blocks(W, H, B)

    def search(self, cur, ref, B, R, begin=0, end=None, nthreads=None):
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        nb = self.num_blocks(W, H, B)
        end = nb if end is None else end
        out = np.zeros(end - begin, RESULT_DTYPE)
        nthreads = nthreads or min(32, os.cpu_count() or 1)
        rc = self.lib.me_oracle_search(pc, pr, W, H, B, R, begin, end, nthreads, out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def search_ssim(self, cur, ref, B, R, begin=0, end=None, nthreads=None):
        """SSIM-cost full search (main_ssim.c / ssim.c); 'ssd' = 1 when a candidate scored > 0."""
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        nb = self.num_blocks(W, H, B)
        end = nb if end is None else end
        out = np.zeros(end - begin, RESULT_DTYPE)
        nthreads = nthreads or min(32, os.cpu_count() or 1)
        rc = self.lib.me_oracle_search_ssim(pc, pr, W, H, B, R, begin, end, nthreads,
                                            out.ctypes.data_as(C.c_void_p))
        assert rc == 0, rc
        return out

    def search_fast(self, cur, ref, B, R, algo, begin=0, end=None, nthreads=None):
        """Three-step (algo 1) / diamond (algo 2) search as DEFINED by oracle/me_oracle_fast.c
        (parity unpinned: the reference has no fast search).  Returns (results, evaluations)."""
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        nb = self.num_blocks(W, H, B)
        end = nb if end is None else end
        out = np.zeros(end - begin, RESULT_DTYPE)
        ev = C.c_uint64(0)
        nthreads = nthreads or min(32, os.cpu_count() or 1)
        rc = self.lib.me_oracle_search_fast(pc, pr, W, H, B, R, algo, begin, end, nthreads,
                                            out.ctypes.data_as(C.c_void_p), C.byref(ev))
        assert rc == 0, rc
        return out, int(ev.value)

    def ssim_frame_scores(self, cur, ref, mc):
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        mc, pm = _u8(mc)
        a, b = C.c_float(0), C.c_float(0)
        self.lib.me_oracle_ssim_frame_scores(pc, pr, pm, cur.size, C.byref(a), C.byref(b))
        return a.value, b.value

    def pixel_compares(self, W, H, B, R):
        return int(self.lib.me_oracle_pixel_compares(W, H, B, R))

    def candidates(self, W, H, B, R):
        return int(self.lib.me_oracle_candidates(W, H, B, R))

    def output5(self, cur, ref, B, res):
        """5 stacked planes of main.c:160-168 + PSNR, from an oracle result array."""
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        res = np.ascontiguousarray(res)
        mc = np.zeros((H, W), np.uint8)
        self.lib.me_oracle_motion_compensate(pr, W, H, B, res.ctypes.data_as(C.c_void_p),
                                             mc.ctypes.data_as(C.c_void_p))
        d0 = np.zeros((H, W), np.uint8)
        d1 = np.zeros((H, W), np.uint8)
        self.lib.me_oracle_frame_diff(pr, pc, W * H, d0.ctypes.data_as(C.c_void_p))
        self.lib.me_oracle_frame_diff(mc.ctypes.data_as(C.c_void_p), pc, W * H, d1.ctypes.data_as(C.c_void_p))
        psnr = self.lib.me_oracle_psnr(mc.ctypes.data_as(C.c_void_p), pc, W, H)
        return np.concatenate([ref, cur, mc, d0, d1], axis=0), psnr


class Ref:
    """The unmodified reference CPU path (findBestBlkMse etc.)."""

    def __init__(self):
        self.lib = C.CDLL(REF_SO)
        L = self.lib
        L.ref_search_blocks.restype = C.c_int
        L.ref_search_blocks.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.ref_search_pool.restype = C.c_double
        L.ref_search_pool.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.ref_postprocess.restype = C.c_double
        L.ref_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]
        L.ref_sizeof_block.restype = C.c_int
        L.ref_sizeof_prediction_frame.restype = C.c_int

    @staticmethod
    def available():
        return os.path.exists(REF_SO)

    def search(self, cur, ref, B, R, begin=0, end=None, nthreads=None):
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        nb = (-(-W // B)) * (-(-H // B))
        end = nb if end is None else end
        out = np.zeros(end - begin, RESULT_DTYPE)
        nthreads = nthreads or min(32, os.cpu_count() or 1)
        rc = self.lib.ref_search_blocks(pc, pr, W, H, B, R, begin, end, nthreads, out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out

    def search_pool(self, cur, ref, B, R, begin=0, end=None, pool_threads=100):
        """The reference's own timed region (main.c:144-158). Returns (seconds, results)."""
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        nb = (-(-W // B)) * (-(-H // B))
        end = nb if end is None else end
        out = np.zeros(end - begin, RESULT_DTYPE)
        sec = self.lib.ref_search_pool(pc, pr, W, H, B, R, begin, end, pool_threads, out.ctypes.data_as(C.c_void_p))
        assert sec >= 0
        return sec, out

    def output5(self, cur, ref, B, mvx, mvy):
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        mvx = np.ascontiguousarray(mvx, np.int32)
        mvy = np.ascontiguousarray(mvy, np.int32)
        out = np.zeros((5 * H, W), np.uint8)
        psnr = self.lib.ref_postprocess(pc, pr, W, H, B, mvx.ctypes.data_as(C.c_void_p),
                                        mvy.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out, psnr


class RefSsim:
    """The unmodified reference SSIM search (findBestBlkSSIM, main_ssim.c:16 / ssim.c)."""

    def __init__(self):
        self.lib = C.CDLL(REF_SSIM_SO)
        L = self.lib
        L.ref_ssim_search_blocks.restype = C.c_int
        L.ref_ssim_search_blocks.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 7 + [C.c_void_p]
        L.ref_ssim_score.restype = C.c_float
        L.ref_ssim_score.argtypes = [C.c_void_p, C.c_void_p] + [C.c_int] * 6
        L.ref_ssim_postprocess.restype = None
        L.ref_ssim_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]

    @staticmethod
    def available():
        return os.path.exists(REF_SSIM_SO)

    def search(self, cur, ref, B, R, begin=0, end=None, nthreads=None):
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        nb = (-(-W // B)) * (-(-H // B))
        end = nb if end is None else end
        out = np.zeros(end - begin, RESULT_DTYPE)
        nthreads = nthreads or min(32, os.cpu_count() or 1)
        rc = self.lib.ref_ssim_search_blocks(pc, pr, W, H, B, R, begin, end, nthreads,
                                             out.ctypes.data_as(C.c_void_p))
        assert rc == 0
        return out

    def output5(self, cur, ref, B, mvx, mvy):
        cur, pc = _u8(cur)
        ref, pr = _u8(ref)
        H, W = cur.shape
        mvx = np.ascontiguousarray(mvx, np.int32)
        mvy = np.ascontiguousarray(mvy, np.int32)
        out = np.zeros((5 * H, W), np.uint8)
        a, b = C.c_float(0), C.c_float(0)
        self.lib.ref_ssim_postprocess(pc, pr, W, H, B, mvx.ctypes.data_as(C.c_void_p),
                                      mvy.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                      C.byref(a), C.byref(b))
        return out, a.value, b.value


def field_sha(mvx, mvy, ssd):
    """First 16 hex of SHA-256 over per-block little-endian (i32 mvx, i32 mvy, u32 ssd)."""
    rec = np.zeros(len(mvx), np.dtype([("x", "<i4"), ("y", "<i4"), ("s", "<u4")]))
    rec["x"], rec["y"], rec["s"] = mvx, mvy, ssd
    return hashlib.sha256(rec.tobytes()).hexdigest()[:16]
