#!/usr/bin/env python3
"""SASS evidence for profiles/: per kernel of libtcs.so, the tensor-core / TMA / TMEM mnemonics and their counts.
    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt        (no GPU needed: cuobjdump reads the .so)"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vae-diffusion-toy-crystals_b200", "toycrystals_b200", "libtcs.so")
PAT = re.compile(r"\b(UTC[A-Z0-9.]*|UTMA[A-Z0-9.]*|LDTM[A-Z0-9.x]*|STTM[A-Z0-9.x]*|HMMA[A-Z0-9.]*|FFMA2|UBLKCP[A-Z0-9.]*)\b")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input=sass, capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in names.splitlines():
        m = re.search(r"Function : (.*)", line)
        if m:
            cur = re.sub(r"\(.*", "", m.group(1).replace("(anonymous namespace)::", "")).strip()
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = PAT.search(line.split("/*")[1] if line.lstrip().startswith("/*") and line.count("/*") > 1 else line)
        if m and "/*" in line:
            per[cur][m.group(1)] += 1
    print("# SASS evidence, libtcs.so (cuobjdump -sass), per kernel: mnemonic x count")
    print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), UTMALDG/UTMASTG/UTMAREDG = TMA load / store / reduce-add,")
    print("# LDTM/STTM = tcgen05.ld/st, HMMA = legacy mma.sync, FFMA2 = packed fp32x2 math, UTCBAR = tcgen05.commit")
    print("# conv_tc_kernel<N, EPI, MSUB, CG, GEO>: GEO 1 = tap-shift geometry (one window per channel block, taps as descriptor shifts)")
    for k, c in per.items():
        if not c:
            continue
        print(f"\n{k}")
        print("    " + "  ".join(f"{m} x{n}" for m, n in sorted(c.items())))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    print("\n# totals: " + "  ".join(f"{m} x{n}" for m, n in sorted(tot.items())))


if __name__ == "__main__":
    sys.exit(main())
