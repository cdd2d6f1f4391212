"""CPU tests of the C-ABI boundary: the library loads without a GPU, exports
every symbol include/*.h declares, keeps the reference's struct layouts, and
fails loudly (no CPU fallback) when there is no device."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import motionestimation_b200 as me
from oracle_binding import Ref, ROOT

INCLUDE = os.path.join(ROOT, "include")


def declared_symbols():
    names = set()
    for h in ("me_b200.h", "me_common.h"):
        src = open(os.path.join(INCLUDE, h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        for m in re.finditer(r"^[A-Za-z_][\w \*]*?\b(\w+)\s*\(", src, flags=re.M):
            name = m.group(1)
            if name not in ("defined", "extern"):
                names.add(name)
    return names


def test_header_symbols_exported():
    lib = me.load_library()
    syms = declared_symbols()
    assert {"me_b200_search", "me_b200_create", "me_b200_submit", "me_b200_search_device",
            "createPredictionFrame", "yuvReadFrame", "motionCompensatedFrame", "imagePSNR"} <= syms
    for s in sorted(syms):
        assert hasattr(lib, s), f"{s} declared in include/ but not exported by libme_b200.so"


def test_exports_are_plain_c():
    out = subprocess.run(["nm", "-D", "--defined-only", me.library_path()], capture_output=True, text=True,
                         check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for s in declared_symbols():
        assert s in exported
    assert me.load_library().me_b200_abi_version() == 3


def test_struct_layouts_match_reference():
    assert C.sizeof(me.Block) == 44            # block.h:6-19, 11 ints
    assert C.sizeof(me.PredictionFrame) == 32  # prediction_frame.h:8-16 on LP64
    if Ref.available():
        r = Ref().lib
        assert r.ref_sizeof_block() == C.sizeof(me.Block)
        assert r.ref_sizeof_prediction_frame() == C.sizeof(me.PredictionFrame)


def test_header_compiles_as_c99(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "me_b200.h"\nint main(void){ me_b200_ctx *c = 0; (void)c; return ME_OK; }\n')
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", INCLUDE, "-c", str(src),
                    "-o", str(tmp_path / "t.o")], check=True)


def test_argument_validation_needs_no_gpu():
    lib = me.load_library()
    h = C.c_void_p()
    assert lib.me_b200_create(C.byref(h), 0, 0, 288, 8, 12) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_create(C.byref(h), 0, 352, 288, 0, 12) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_create(C.byref(h), 0, 352, 288, 8, -1) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_create(None, 0, 352, 288, 8, 12) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_create_ex(C.byref(h), 0, 352, 288, 8, 12, 0, 0) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_create_ex(C.byref(h), 0, 352, 288, 8, 12, 1, 9) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_create(C.byref(h), 0, 352, 288, 300, 12) == me.ME_ERR_UNSUPPORTED
    assert lib.me_b200_search(None, None, 12) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_wait(None, 0) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_num_blocks(None) == 0
    assert lib.me_b200_strerror(me.ME_ERR_NO_DEVICE).decode().startswith("no usable CUDA device")
    lib.me_b200_destroy(None)  # no-op
    # peer-field entry points
    assert lib.me_b200_device_alloc(None, 16) is None
    assert lib.me_b200_peer_barrier(None, None, 2, 0, 1, 10, None) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_search_device_band_peers(None, None, None, 0, 0, 1, 0, 0, None, None, 0, None) == \
        me.ME_ERR_INVALID_ARG
    assert lib.me_b200_ipc_export(None, None, None) == me.ME_ERR_INVALID_ARG
    assert C.sizeof(me.Field) == 32


@pytest.mark.skipif(me.device_count() > 0, reason="this test is about machines without a GPU")
def test_no_cpu_fallback_without_gpu():
    """Without a device every compute entry point reports ME_ERR_NO_DEVICE."""
    with pytest.raises(me.MeError) as ei:
        me.Estimator(352, 288, 8, 12)
    assert ei.value.code == me.ME_ERR_NO_DEVICE
    cur = np.zeros(352 * 288, np.int32)
    pf = me.create_prediction_frame(cur, 352, 288, 8)
    with pytest.raises(me.MeError) as ei:
        me.search_prediction_frame(pf, cur.copy(), 12)
    assert ei.value.code == me.ME_ERR_NO_DEVICE
    assert pf.blks[0].is_best_match_found == 0
    rate, _ = me.int_peak(0)
    assert rate == 0.0


@pytest.mark.skipif(me.device_count() > 0, reason="no-GPU behaviour")
def test_cli_fails_loudly_without_gpu(tmp_path):
    exe = os.path.join(ROOT, "motionestimation_b200", "mes_b200")
    g = os.path.join(ROOT, "motionestimation_b200", "data")   # the reference's frames/ directory
    p = subprocess.run([exe, f"{g}/ForemanYF4.yuv", f"{g}/ForemanYF1.yuv", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 2 and "no usable CUDA device" in p.stderr
    assert not os.path.exists(tmp_path / "output_8_12.yuv")
    # the SSIM program and the fast patterns have no CPU path either
    exe = os.path.join(ROOT, "motionestimation_b200", "mes_b200_ssim")
    p = subprocess.run([exe, f"{g}/ForemanYF4.yuv", f"{g}/ForemanYF1.yuv", str(tmp_path), "16", "7", "352", "288"],
                       capture_output=True, text=True)
    assert p.returncode == 2 and "no usable CUDA device" in p.stderr
    assert not os.path.exists(tmp_path / "output_16_7.yuv")


@pytest.mark.skipif(me.device_count() > 0, reason="no-GPU behaviour")
def test_ssim_and_fast_entry_points_need_a_device():
    lib = me.load_library()
    cur = np.zeros(64 * 48, np.int32)
    pf = me.create_prediction_frame(cur, 64, 48, 8)
    refp = cur.ctypes.data_as(C.POINTER(C.c_int))
    assert lib.me_b200_search_ssim(C.byref(pf), refp, 4) == me.ME_ERR_NO_DEVICE
    assert lib.me_b200_search_fast(C.byref(pf), refp, 4, me.ME_SEARCH_DIAMOND, None, None) == me.ME_ERR_NO_DEVICE
    assert lib.me_b200_search_fast(C.byref(pf), refp, 4, 0, None, None) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_set_cost(None, me.ME_COST_SSIM) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_set_search(None, me.ME_SEARCH_DIAMOND) == me.ME_ERR_INVALID_ARG
    assert lib.me_b200_tss_first_step(7) == 4 and lib.me_b200_tss_first_step(32) == 16


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under motionestimation_b200/ or
    include/ may reference it."""
    bad = []
    for base in ("motionestimation_b200", "include"):
        for dp, _, fns in os.walk(os.path.join(ROOT, base)):
            if "_build" in dp or "__pycache__" in dp:
                continue
            for fn in fns:
                if fn.endswith((".py", ".c", ".cu", ".cuh", ".h", "Makefile")):
                    txt = open(os.path.join(dp, fn), errors="ignore").read()
                    if re.search(r"me_oracle|oracle/|oracle_binding|libme_ref|/root/reference", txt):
                        bad.append(os.path.join(dp, fn))
    assert not bad, bad
