"""INTEGRATION.md section 1 is tested, not just documented: the diff printed there is applied to the
unmodified reference src/cpu/main.c and built with the documented gcc line against libme_b200.so
(oracle/build_patched_ref.py -> oracle/_ref/mes_ref_patched; built where /root/reference exists,
shipped prebuilt to the GPU box).  The reference's own main(), yuvReadFrame, createPredictionFrame,
motionCompensatedFrame, frameDiff, imagePSNR and yuvWriteFrame then run around ONE me_b200_search
call, and the output file must be the reference's shipped golden."""
import hashlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATCHED = os.path.join(ROOT, "oracle", "_ref", "mes_ref_patched")
FRAMES = os.path.join(ROOT, "motionestimation_b200", "data")
HAVE_REF = os.path.isdir("/root/reference/src/cpu")


@pytest.mark.skipif(not HAVE_REF, reason="needs the reference checkout (build container only)")
def test_documented_patch_applies_and_builds(tmp_path):
    out = tmp_path / "mes_patched"
    p = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "build_patched_ref.py"), "--out", str(out)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert out.exists()
    # the thread pool is gone from the link, the search symbol comes from the shared library
    syms = subprocess.run(["nm", "-D", str(out)], capture_output=True, text=True).stdout
    assert " U me_b200_search" in syms and "thpool" not in syms
    # the program's own src/common functions are the ones linked in (defined, not imported)
    full = subprocess.run(["nm", str(out)], capture_output=True, text=True).stdout
    for name in ("createPredictionFrame", "yuvReadFrame", "motionCompensatedFrame", "imagePSNR"):
        assert f" T {name}" in full, name


def test_patched_reference_without_gpu_fails_loudly(tmp_path):
    """No CPU fallback: on a box without a GPU the patched reference reports the error and exits 2."""
    import motionestimation_b200 as me
    if not os.path.exists(PATCHED):
        pytest.skip("oracle/_ref/mes_ref_patched not built")
    if me.device_count() > 0:
        pytest.skip("a GPU is present")
    p = subprocess.run([PATCHED, f"{FRAMES}/ForemanYF4.yuv", f"{FRAMES}/ForemanYF1.yuv", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 2 and "no usable CUDA device" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("args,md5", [(["4", "15"], "686f3f74e7dc7f2e8321b513eed033e8"),
                                      (["4", "7"], "a77c2741268fc18e5f73c599fc040172")])
def test_patched_reference_reproduces_the_shipped_goldens(tmp_path, args, md5):
    """results/cpu/foreman/output_4_15.yuv / output_4_7.yuv (cur = YF4, ref = YF1; 4_15.txt:2-8)."""
    assert os.path.exists(PATCHED), "oracle/_ref/mes_ref_patched must be shipped prebuilt"
    p = subprocess.run([PATCHED, f"{FRAMES}/ForemanYF4.yuv", f"{FRAMES}/ForemanYF1.yuv", str(tmp_path), *args],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    data = (tmp_path / f"output_{args[0]}_{args[1]}.yuv").read_bytes()
    assert hashlib.md5(data).hexdigest() == md5


@pytest.mark.gpu
def test_patched_reference_default_run_prints_the_logged_psnr(tmp_path):
    """results/cpu/foreman/2990wx_threadripper_64_cores.txt:10 -- PSNR: 31.816000 (defaults 8 / 12)."""
    assert os.path.exists(PATCHED)
    p = subprocess.run([PATCHED, f"{FRAMES}/ForemanYF4.yuv", f"{FRAMES}/ForemanYF1.yuv", str(tmp_path)],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    assert "PSNR: 31.816000" in p.stdout
