#!/usr/bin/env python
"""oracle/build_patched_ref.py -- TEST INFRASTRUCTURE: proves INTEGRATION.md section 1.

Takes the unified diff printed in INTEGRATION.md section 1 (the first ```diff block), applies it to a
scratch copy of the UNMODIFIED reference `src/cpu/main.c` (outside the repository; the reference's
`../common` is reached through a symlink, nothing is copied into the repo), and builds the result
with the gcc line of the same section against `libme_b200.so`: the reference's own main(), I/O and
post-processing around one `me_b200_search` call.  Output: oracle/_ref/mes_ref_patched (git-ignored,
shipped to the GPU box prebuilt like the rest of oracle/_ref).

    python oracle/build_patched_ref.py [--ref /root/reference] [--out oracle/_ref/mes_ref_patched]

The reference's sources have CRLF line endings, so the diff (kept with LF in the markdown) is
converted to CRLF and applied with `patch --binary`, as the document tells a maintainer to do.
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def extract_diff(md_path):
    text = open(md_path, encoding="utf-8").read()
    m = re.search(r"```diff\n(--- a/src/cpu/main\.c\n.*?)```", text, re.S)
    if not m:
        raise SystemExit("no ```diff block for src/cpu/main.c in %s" % md_path)
    return m.group(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_ref", "mes_ref_patched"))
    ap.add_argument("--keep", action="store_true", help="print the scratch directory instead of deleting it")
    args = ap.parse_args()
    src_main = os.path.join(args.ref, "src", "cpu", "main.c")
    if not os.path.exists(src_main):
        raise SystemExit("reference not found at %s" % args.ref)
    libdir = os.path.join(ROOT, "motionestimation_b200")
    if not os.path.exists(os.path.join(libdir, "libme_b200.so")):
        raise SystemExit("build motionestimation_b200/libme_b200.so first")
    diff = extract_diff(os.path.join(ROOT, "INTEGRATION.md"))
    scratch = tempfile.mkdtemp(prefix="me_b200_integration_")
    try:
        os.makedirs(os.path.join(scratch, "src", "cpu"))
        os.symlink(os.path.join(args.ref, "src", "common"), os.path.join(scratch, "src", "common"))
        shutil.copy(src_main, os.path.join(scratch, "src", "cpu", "main.c"))
        crlf = diff.replace("\r\n", "\n").replace("\n", "\r\n").encode()
        subprocess.run(["patch", "--binary", "-p1"], input=crlf, cwd=scratch, check=True)
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        out = os.path.abspath(args.out)
        rel = os.path.relpath(libdir, os.path.dirname(out))
        # the build line of INTEGRATION.md section 1 (run from src/cpu); rpath relative to the binary so the
        # pair travels together
        cmd = ["gcc", "../common/block.c", "../common/prediction_frame.c", "../common/utils.c", "main.c",
               "-I" + os.path.join(ROOT, "include"), "-L" + libdir, "-lme_b200",
               "-Wl,-rpath,$ORIGIN/" + rel, "-o", out, "-lm", "-w"]
        subprocess.run(cmd, cwd=os.path.join(scratch, "src", "cpu"), check=True)
        print("built", out)
    finally:
        if args.keep:
            print("scratch:", scratch)
        else:
            shutil.rmtree(scratch, ignore_errors=True)


if __name__ == "__main__":
    sys.exit(main())
