"""ctypes binding of the C ABI (``include/me_b200.h``) and the C host layer
(``include/me_common.h``).  Names follow the reference's ``src/common`` interface:
``Block`` = ``block`` (block.h:6-19), ``PredictionFrame`` = ``predictionFrame``
(prediction_frame.h:8-16), ``create_prediction_frame`` = ``createPredictionFrame``
(prediction_frame.c:3-25), ``search_prediction_frame`` = the dispatch loop of
``src/cpu/main.c:144-158``.

Device memory, streams and process groups come from torch (plumbing); the
search itself always runs in ``libme_b200.so``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

ME_OK = 0
ME_ERR_INVALID_ARG = -1
ME_ERR_UNSUPPORTED = -2
ME_ERR_CUDA = -3
ME_ERR_NO_DEVICE = -4
ME_ERR_NOMEM = -5
ME_ERR_STATE = -6
ME_KERNEL_AUTO, ME_KERNEL_GENERIC, ME_KERNEL_TILED, ME_KERNEL_DIRECT = 0, 1, 2, 3
ME_KERNEL_SSIM, ME_KERNEL_FAST = 4, 5
ME_HOST_WRITE_COMBINED = 1
ME_COST_MSE, ME_COST_SSIM = 0, 1
ME_SEARCH_FULL, ME_SEARCH_THREE_STEP, ME_SEARCH_DIAMOND = 0, 1, 2
ME_B200_MAX_SLOTS = 4
ME_B200_MAX_PEERS = 8
ME_B200_IPC_HANDLE_BYTES = 64

PEAK_NAMES = ["IDP4A", "VABSDIFF4", "SSD_PAIR", "IADD3", "LOP3", "IMAD", "VIMNMX",
              "SSD_PAIR_LDS", "IDP4A_IADD3", "LOOP_REPLICA"]


class MeError(RuntimeError):
    def __init__(self, code: int, what: str, detail: str = ""):
        self.code = code
        super().__init__(f"{what}: error {code} ({_strerror(code)}) {detail}".strip())


class Block(C.Structure):
    """``block`` -- src/common/block.h:6-19 (11 ints, 44 bytes)."""
    _fields_ = [(n, C.c_int) for n in (
        "idx_x", "idx_y", "top_left_x", "top_left_y", "bottom_right_x", "bottom_right_y",
        "width", "height", "is_best_match_found", "motion_vectorX", "motion_vectorY")]


class Field(C.Structure):
    """``me_b200_field``: device pointers of one copy of the motion field (any may be NULL)."""
    _fields_ = [("mvx", C.c_void_p), ("mvy", C.c_void_p), ("ssd", C.c_void_p), ("score", C.c_void_p)]


class PredictionFrame(C.Structure):
    """``predictionFrame`` -- src/common/prediction_frame.h:8-16."""
    _fields_ = [("frame", C.POINTER(C.c_int)), ("width", C.c_int), ("height", C.c_int),
                ("blk_dim", C.c_int), ("num_blks", C.c_int), ("blks", C.POINTER(Block))]


_LIB: Optional[C.CDLL] = None


def library_path() -> str:
    # ME_B200_LIBRARY: another build of the same library (A/B experiments with tools/quick_bench.py)
    return os.environ.get("ME_B200_LIBRARY") or os.path.join(os.path.dirname(os.path.abspath(__file__)),
                                                             "libme_b200.so")


def load_library() -> C.CDLL:
    """Load ``libme_b200.so``; raises (loudly) if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `make -C motionestimation_b200` "
            "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no fallback.")
    lib = C.CDLL(path)
    vp, i32p, u32p, f32p, u8p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    sig = {
        "me_b200_abi_version": (C.c_int, []),
        "me_b200_strerror": (C.c_char_p, [C.c_int]),
        "me_b200_last_error": (C.c_char_p, [vp]),
        "me_b200_device_count": (C.c_int, []),
        "me_b200_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
        "me_b200_create_ex": (C.c_int, [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int]),
        "me_b200_destroy": (None, [vp]),
        "me_b200_num_blocks": (C.c_int, [vp]),
        "me_b200_blocks_x": (C.c_int, [vp]),
        "me_b200_blocks_y": (C.c_int, [vp]),
        "me_b200_kernel_in_use": (C.c_int, [vp]),
        "me_b200_last_kernel": (C.c_int, [vp]),
        "me_b200_fallback_launches": (C.c_uint64, [vp]),
        "me_b200_pixel_compares": (C.c_uint64, [vp]),
        "me_b200_candidates": (C.c_uint64, [vp]),
        "me_b200_launch_count": (C.c_uint64, [vp]),
        "me_b200_search": (C.c_int, [C.POINTER(PredictionFrame), C.POINTER(C.c_int), C.c_int]),
        "me_b200_search_scores": (C.c_int, [C.POINTER(PredictionFrame), C.POINTER(C.c_int), C.c_int,
                                            f32p, u32p]),
        "me_b200_release_cached": (None, []),
        "me_b200_set_cost": (C.c_int, [vp, C.c_int]),
        "me_b200_set_search": (C.c_int, [vp, C.c_int]),
        "me_b200_fast_evaluations": (C.c_int, [vp, C.POINTER(C.c_uint64)]),
        "me_b200_tss_first_step": (C.c_int, [C.c_int]),
        "me_b200_search_ssim": (C.c_int, [C.POINTER(PredictionFrame), C.POINTER(C.c_int), C.c_int]),
        "me_b200_search_ssim_scores": (C.c_int, [C.POINTER(PredictionFrame), C.POINTER(C.c_int), C.c_int,
                                                 f32p, u32p]),
        "me_b200_search_fast": (C.c_int, [C.POINTER(PredictionFrame), C.POINTER(C.c_int), C.c_int, C.c_int,
                                          f32p, u32p]),
        "me_b200_search_u8": (C.c_int, [vp, u8p, u8p, C.c_int, i32p, i32p, u32p, f32p]),
        "me_b200_submit": (C.c_int, [vp, C.c_int, u8p, u8p, C.c_int, i32p, i32p, u32p, f32p]),
        "me_b200_wait": (C.c_int, [vp, C.c_int]),
        "me_b200_set_ingest_helper": (C.c_int, [vp, C.c_int, C.c_int]),
        "me_b200_submit_sequence": (C.c_int, [vp, C.c_int, u8p, C.c_int, i32p, i32p, u32p, f32p]),
        "me_b200_search_sequence_u8": (C.c_int, [vp, u8p, C.c_int, i32p, i32p, u32p, f32p]),
        "me_b200_host_alloc": (vp, [C.c_size_t]),
        "me_b200_host_alloc_ex": (vp, [C.c_size_t, C.c_int]),
        "me_b200_host_free": (None, [vp]),
        "me_b200_search_device": (C.c_int, [vp, u8p, u8p, C.c_size_t, C.c_size_t, C.c_int,
                                            i32p, i32p, u32p, f32p, vp]),
        "me_b200_search_device_band": (C.c_int, [vp, u8p, u8p, C.c_size_t, C.c_size_t, C.c_int,
                                                 C.c_int, C.c_int, i32p, i32p, u32p, f32p, vp]),
        "me_b200_device_alloc": (vp, [vp, C.c_size_t]),
        "me_b200_device_free": (None, [vp, vp]),
        "me_b200_ipc_export": (C.c_int, [vp, vp, C.c_char_p]),
        "me_b200_ipc_open": (C.c_int, [vp, C.c_char_p, C.POINTER(vp)]),
        "me_b200_ipc_close": (C.c_int, [vp, vp]),
        "me_b200_search_device_band_peers": (C.c_int, [vp, u8p, u8p, C.c_size_t, C.c_size_t, C.c_int, C.c_int,
                                                       C.c_int, C.POINTER(Field), C.POINTER(Field), C.c_int, vp]),
        "me_b200_peer_barrier": (C.c_int, [vp, C.POINTER(vp), C.c_int, C.c_int, C.c_uint32, C.c_int, vp]),
        "me_b200_peer_barrier_timed_out": (C.c_int, [vp, C.POINTER(C.c_int)]),
        "me_b200_postprocess_device": (C.c_int, [vp, u8p, u8p, C.c_size_t, i32p, i32p, u8p, vp, vp, vp]),
        "me_b200_postprocess_device_batch": (C.c_int, [vp, u8p, u8p, C.c_size_t, C.c_size_t, C.c_int, i32p, i32p, u8p,
                                                       C.c_size_t, vp, vp, vp]),
        "me_b200_int_peak": (C.c_double, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]),
        # host layer (include/me_common.h)
        "createBlk": (None, [C.POINTER(Block)] + [C.c_int] * 6),
        "createPredictionFrame": (None, [C.POINTER(PredictionFrame), C.POINTER(C.c_int), C.c_int,
                                         C.c_int, C.c_int]),
        "getTimeStamp": (C.c_double, []),
        "yuvReadFrame": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int]),
        "yuvReadFrameU8": (C.c_int, [C.c_char_p, vp, C.c_int]),
        "yuvWriteFrame": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.c_int]),
        "frameDiff": (None, [C.POINTER(C.c_int)] * 3 + [C.c_int]),
        "motionCompensatedFrame": (C.c_int, [C.POINTER(C.c_int), PredictionFrame, C.POINTER(C.c_int)]),
        "imagePSNR": (C.c_double, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int, C.c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


EXPORTED_SYMBOLS = None  # filled lazily by tests from include/*.h


def _strerror(code: int) -> str:
    try:
        return load_library().me_b200_strerror(code).decode()
    except Exception:  # pragma: no cover
        return "?"


def device_count() -> int:
    return load_library().me_b200_device_count()


def _np_ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Estimator:
    """One search context (``me_b200_ctx``): one GPU, one geometry.

    Replaces the per-run setup of ``src/cpu/main.c:117-143`` and, through
    :meth:`search_u8` / :meth:`search_device`, the timed dispatch loop
    ``main.c:144-158``.
    """

    def __init__(self, width: int, height: int, blk_dim: int = 8, extra_span: int = 12,
                 device: int = 0, max_pairs: int = 1, kernel: int = ME_KERNEL_AUTO,
                 cost: int = ME_COST_MSE, search: int = ME_SEARCH_FULL):
        self._lib = load_library()
        h = C.c_void_p()
        rc = self._lib.me_b200_create_ex(C.byref(h), device, width, height, blk_dim, extra_span,
                                         max_pairs, kernel)
        if rc != ME_OK:
            raise MeError(rc, "me_b200_create_ex", self._lib.me_b200_last_error(None).decode())
        self._h = h
        self.width, self.height, self.blk_dim, self.extra_span = width, height, blk_dim, extra_span
        self.device, self.max_pairs = device, max_pairs
        self.num_blocks = self._lib.me_b200_num_blocks(h)
        self.blocks_x = self._lib.me_b200_blocks_x(h)
        self.blocks_y = self._lib.me_b200_blocks_y(h)
        self.cost, self.search = cost, search
        if cost != ME_COST_MSE:
            self._check(self._lib.me_b200_set_cost(h, cost), "me_b200_set_cost")
        if search != ME_SEARCH_FULL:
            self._check(self._lib.me_b200_set_search(h, search), "me_b200_set_search")

    @property
    def fast_evaluations(self) -> int:
        """Candidate evaluations of the fast searches run so far (synchronises the device)."""
        v = C.c_uint64(0)
        self._check(self._lib.me_b200_fast_evaluations(self._h, C.byref(v)), "me_b200_fast_evaluations")
        return int(v.value)

    # -- bookkeeping -----------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.me_b200_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def kernel_in_use(self) -> int:
        return self._lib.me_b200_kernel_in_use(self._h)

    @property
    def last_kernel(self) -> int:
        return self._lib.me_b200_last_kernel(self._h)

    @property
    def fallback_launches(self) -> int:
        """AUTO launches the tuned kernel could not take (the generic kernel ran): 0 when all is well."""
        return int(self._lib.me_b200_fallback_launches(self._h))

    @property
    def pixel_compares(self) -> int:
        return int(self._lib.me_b200_pixel_compares(self._h))

    @property
    def candidates(self) -> int:
        return int(self._lib.me_b200_candidates(self._h))

    @property
    def launch_count(self) -> int:
        return int(self._lib.me_b200_launch_count(self._h))

    def _check(self, rc: int, what: str):
        if rc != ME_OK:
            raise MeError(rc, what, self._lib.me_b200_last_error(self._h).decode())

    # -- host u8 path -------------------------------------------------------------------
    def _host_args(self, cur, ref):
        cur = np.ascontiguousarray(cur, dtype=np.uint8)
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        n = self.width * self.height
        if cur.size != ref.size or cur.size % n:
            raise ValueError("frames must be npairs x H x W uint8")
        return cur, ref, cur.size // n

    def search_u8(self, cur: np.ndarray, ref: np.ndarray):
        """Blocking host-buffer search. Returns dict of (npairs, num_blocks) arrays."""
        cur, ref, npairs = self._host_args(cur, ref)
        out = self._alloc_out(npairs)
        self._check(self._lib.me_b200_search_u8(self._h, _np_ptr(cur), _np_ptr(ref), npairs,
                                                _np_ptr(out["mvx"]), _np_ptr(out["mvy"]),
                                                _np_ptr(out["ssd"]), _np_ptr(out["score"])),
                    "me_b200_search_u8")
        return out

    def search_sequence_u8(self, frames: np.ndarray):
        """Consecutive pairs of a frame sequence (nframes, H, W): pair i = frame i+1 searched in
        frame i; every frame is uploaded once.  Returns dict of (nframes-1, num_blocks) arrays."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        n = self.width * self.height
        if frames.size % n or frames.size // n < 2:
            raise ValueError("frames must be nframes x H x W uint8, nframes >= 2")
        nframes = frames.size // n
        out = self._alloc_out(nframes - 1)
        self._check(self._lib.me_b200_search_sequence_u8(self._h, _np_ptr(frames), nframes,
                                                         _np_ptr(out["mvx"]), _np_ptr(out["mvy"]),
                                                         _np_ptr(out["ssd"]), _np_ptr(out["score"])),
                    "me_b200_search_sequence_u8")
        return out

    def submit_sequence_ptr(self, slot: int, frames_ptr: int, nframes: int, mvx_ptr: int, mvy_ptr: int,
                            ssd_ptr: int, score_ptr: int):
        self._check(self._lib.me_b200_submit_sequence(self._h, slot, frames_ptr, nframes, mvx_ptr, mvy_ptr,
                                                      ssd_ptr or None, score_ptr or None),
                    "me_b200_submit_sequence")

    def _alloc_out(self, npairs: int):
        nb = self.num_blocks
        return {"mvx": np.empty((npairs, nb), np.int32), "mvy": np.empty((npairs, nb), np.int32),
                "ssd": np.empty((npairs, nb), np.uint32), "score": np.empty((npairs, nb), np.float32)}

    def submit_ptr(self, slot: int, cur_ptr: int, ref_ptr: int, npairs: int, mvx_ptr: int,
                   mvy_ptr: int, ssd_ptr: int, score_ptr: int):
        """Pipelined submit on raw (pinned) host pointers; pair with :meth:`wait`."""
        self._check(self._lib.me_b200_submit(self._h, slot, cur_ptr, ref_ptr, npairs, mvx_ptr, mvy_ptr,
                                             ssd_ptr or None, score_ptr or None), "me_b200_submit")

    def wait(self, slot: int):
        self._check(self._lib.me_b200_wait(self._h, slot), "me_b200_wait")

    def set_ingest_helper(self, helper_device: int, helper_pairs: int):
        """The last `helper_pairs` pairs of every submit travel over `helper_device`'s host link and NVLink."""
        self._check(self._lib.me_b200_set_ingest_helper(self._h, helper_device, helper_pairs),
                    "me_b200_set_ingest_helper")

    # -- device-resident path (torch tensors provide the memory) ---------------------------
    def search_device(self, d_cur, d_ref, pitch: int, pair_stride: int, npairs: int,
                      d_mvx, d_mvy, d_ssd=None, d_score=None, stream: int = 0,
                      by_begin: int = 0, by_end: Optional[int] = None):
        """Enqueue the search on device pointers (ints or torch tensors)."""
        def p(x):
            if x is None:
                return None
            return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)
        by_end = self.blocks_y if by_end is None else by_end
        self._check(self._lib.me_b200_search_device_band(
            self._h, p(d_cur), p(d_ref), pitch, pair_stride, npairs, by_begin, by_end,
            p(d_mvx), p(d_mvy), p(d_ssd), p(d_score), stream or None), "me_b200_search_device_band")

    # -- band sharding over peer-mapped fields (one process per GPU, CUDA IPC over NVLink) -------
    def device_alloc(self, nbytes: int) -> int:
        p = self._lib.me_b200_device_alloc(self._h, nbytes)
        if not p:
            raise MeError(ME_ERR_NOMEM, "me_b200_device_alloc")
        return int(p)

    def device_free(self, ptr: int):
        self._lib.me_b200_device_free(self._h, ptr)

    def ipc_export(self, ptr: int) -> bytes:
        buf = C.create_string_buffer(ME_B200_IPC_HANDLE_BYTES)
        self._check(self._lib.me_b200_ipc_export(self._h, ptr, buf), "me_b200_ipc_export")
        return buf.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._check(self._lib.me_b200_ipc_open(self._h, handle, C.byref(p)), "me_b200_ipc_open")
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._check(self._lib.me_b200_ipc_close(self._h, ptr), "me_b200_ipc_close")

    def search_device_band_peers(self, d_cur, d_ref, pitch: int, pair_stride: int, npairs: int, by_begin: int,
                                 by_end: int, local: "Field", peers, stream: int = 0):
        """Search block rows [by_begin, by_end) and store the results into `local` and every field of
        `peers` (peer-mapped device memory)."""
        def p(x):
            return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)
        arr = (Field * max(1, len(peers)))(*peers)
        self._check(self._lib.me_b200_search_device_band_peers(
            self._h, p(d_cur), p(d_ref), pitch, pair_stride, npairs, by_begin, by_end, C.byref(local), arr,
            len(peers), stream or None), "me_b200_search_device_band_peers")

    def peer_barrier(self, flag_ptrs, my_rank: int, epoch: int, timeout_ms: int = 2000, stream: int = 0):
        arr = (C.c_void_p * len(flag_ptrs))(*flag_ptrs)
        self._check(self._lib.me_b200_peer_barrier(self._h, arr, len(flag_ptrs), my_rank, epoch, timeout_ms,
                                                   stream or None), "me_b200_peer_barrier")

    def peer_barrier_timed_out(self) -> bool:
        v = C.c_int(0)
        self._check(self._lib.me_b200_peer_barrier_timed_out(self._h, C.byref(v)), "me_b200_peer_barrier_timed_out")
        return bool(v.value)

    def postprocess_device(self, d_cur, d_ref, pitch: int, d_mvx, d_mvy, d_out5, d_sq_err=None,
                           d_max=None, stream: int = 0):
        def p(x):
            if x is None:
                return None
            return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)
        self._check(self._lib.me_b200_postprocess_device(
            self._h, p(d_cur), p(d_ref), pitch, p(d_mvx), p(d_mvy), p(d_out5), p(d_sq_err), p(d_max),
            stream or None), "me_b200_postprocess_device")


    def postprocess_device_batch(self, d_cur, d_ref, pitch: int, pair_stride: int, npairs: int, d_mvx, d_mvy,
                                 d_out5, out_pair_stride: int, d_sq_err=None, d_max=None, stream: int = 0):
        """Post-search stage of a whole batch (main.c:160-171 per pair): 5 stacked planes + the PSNR integers."""
        def p(x):
            if x is None:
                return None
            return int(x.data_ptr()) if hasattr(x, "data_ptr") else int(x)
        self._check(self._lib.me_b200_postprocess_device_batch(
            self._h, p(d_cur), p(d_ref), pitch, pair_stride, npairs, p(d_mvx), p(d_mvy), p(d_out5),
            out_pair_stride, p(d_sq_err), p(d_max), stream or None), "me_b200_postprocess_device_batch")


def create_prediction_frame(cur_int: np.ndarray, width: int, height: int, blk_dim: int) -> PredictionFrame:
    """``createPredictionFrame`` (prediction_frame.c:3-25) on an int32 frame.
    The returned struct borrows ``cur_int``'s memory; keep the array alive."""
    lib = load_library()
    assert cur_int.dtype == np.int32 and cur_int.size == width * height and cur_int.flags.c_contiguous
    pf = PredictionFrame()
    lib.createPredictionFrame(C.byref(pf), cur_int.ctypes.data_as(C.POINTER(C.c_int)), width, height, blk_dim)
    pf._keepalive = cur_int
    return pf


def search_prediction_frame(pf: PredictionFrame, ref_int: np.ndarray, extra_span: int,
                            want_scores: bool = False, cost: int = ME_COST_MSE,
                            search: int = ME_SEARCH_FULL):
    """The drop-in for ``main.c:144-158`` (``cost=ME_COST_SSIM``: for the loop
    ``main_ssim.c:67-77``; ``search`` = a fast pattern: see ``me_b200_search_fast``): fills
    every block of ``pf``.  Returns (scores, ssd) arrays when ``want_scores``."""
    lib = load_library()
    assert ref_int.dtype == np.int32 and ref_int.flags.c_contiguous
    refp = ref_int.ctypes.data_as(C.POINTER(C.c_int))
    sc = np.empty(pf.num_blks, np.float32) if want_scores else None
    sd = np.empty(pf.num_blks, np.uint32) if want_scores else None
    if search != ME_SEARCH_FULL:
        rc = lib.me_b200_search_fast(C.byref(pf), refp, extra_span, search, _np_ptr(sc), _np_ptr(sd))
    elif cost == ME_COST_SSIM:
        rc = lib.me_b200_search_ssim_scores(C.byref(pf), refp, extra_span, _np_ptr(sc), _np_ptr(sd))
    elif want_scores:
        rc = lib.me_b200_search_scores(C.byref(pf), refp, extra_span, _np_ptr(sc), _np_ptr(sd))
    else:
        rc = lib.me_b200_search(C.byref(pf), refp, extra_span)
    if rc != ME_OK:
        raise MeError(rc, "me_b200_search", lib.me_b200_last_error(None).decode())
    return sc, sd


def int_peak(which: int, iters: int = 2000, device: int = 0):
    """Integer-pipe microbenchmark: (lane-instructions/s, SM MHz during the run)."""
    lib = load_library()
    mhz = C.c_double(0.0)
    rate = lib.me_b200_int_peak(device, which, iters, C.byref(mhz))
    return rate, mhz.value
