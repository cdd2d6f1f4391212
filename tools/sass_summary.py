"""python tools/sass_summary.py > profiles/sass_summary.txt

Static SASS evidence of the built library (no GPU needed): for every kernel in
motionestimation_b200/libme_b200.so the target architecture and how often the mnemonics that prove the
design occur -- UTMALDG (TMA tile loads), SYNCS (mbarrier), IDP.4A / VABSDIFF4 (packed-byte integer
pipes), CREDUX / MATCH / SHFL (warp reductions), ATOMS (shared atomics) -- plus registers and total
instruction count.  Tensor-core mnemonics (UTC*MMA, HMMA, IMMA) are listed to show they are absent:
north_star keeps this path off the tensor cores."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "motionestimation_b200", "libme_b200.so")
OPS = ["UTMALDG", "SYNCS", "IDP.4A", "VABSDIFF4", "CREDUX", "MATCH", "SHFL", "ATOMS", "LDS", "LDG", "STG", "SHF", "PRMT",
       "VIMNMX", "VIADDMNMX", "IMAD", "FFMA", "MUFU", "BAR", "UTCHMMA", "UTCIMMA", "HMMA", "IMMA"]


def demangle(names):
    try:
        out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for o in OPS:
                if op == o or op.startswith(o + ".") or (o == "IDP.4A" and op.startswith("IDP.4A")):
                    kernels[cur][o] += 1
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    regs = {}
    fn = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            fn = m.group(1)
        m = re.search(r"REG:(\d+).*?SHARED:(\d+)", line)
        if m and fn:
            regs[fn] = (int(m.group(1)), int(m.group(2)))
    names = demangle(list(kernels))
    print("# SASS summary of motionestimation_b200/libme_b200.so  (cuobjdump -sass, tools/sass_summary.py)")
    print("# architectures in the fat binary: %s" % ", ".join(archs))
    tot = collections.Counter()
    for k, c in kernels.items():
        tot.update(c)
    print("# whole library: " + ", ".join("%s %d" % (o, tot[o]) for o in OPS if tot[o] or o in ("UTCHMMA", "UTCIMMA", "HMMA", "IMMA")))
    print()
    for k, c in kernels.items():
        short = names[k].replace("(anonymous namespace)::", "").replace("void ", "")
        short = re.sub(r"\(.*", "", short)
        r = regs.get(k)
        print("%s" % short)
        print("    instructions %d%s" % (c["_total"], (", registers %d, static smem %d B" % r) if r else ""))
        print("    " + ", ".join("%s %d" % (o, c[o]) for o in OPS if c[o]))
    return 0


if __name__ == "__main__":
    sys.exit(main())
