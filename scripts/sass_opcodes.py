#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove (or disprove) the Blackwell-native claims, from the built
libssdr_b200.so (cuobjdump -sass; runs without a GPU).  Output committed as profiles/r<round>_sass_opcodes.txt.

    python scripts/sass_opcodes.py [path/to/libssdr_b200.so] > profiles/r2_sass_opcodes.txt

tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, tcgen05.cp -> UTCCP, tcgen05.commit -> UTCBAR, TMA tensor copies ->
UTMALDG/UTMASTG, bulk copies -> UBLKCP, bulk L2 prefetch -> UBLKPF, mbarrier -> SYNCS, legacy mma.sync -> HMMA
(/opt/skills/guides/B200_PROFILING.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCBAR", "UTCCP", "UTCATOM", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UBLKPF",
         "SYNCS", "HMMA", "FFMA2", "FADD2", "FMUL2", "FFMA", "MUFU", "LDS", "STS", "LDG", "STG", "ATOMS", "REDUX", "SHFL", "BAR"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "supersdr_b200", "libssdr_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            per[cur][m.group(1)] += 1
            per[cur]["_total"] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
    print("# static SASS instruction counts per kernel, %s" % os.path.relpath(lib, ROOT))
    print("# only non-zero columns of: " + " ".join(WATCH))
    tot = collections.Counter()
    for (name, c), nice in zip(per.items(), demangle):
        nice = nice.replace("(anonymous namespace)::", "").replace("ssdr::", "").replace("void ", "")
        nice = re.sub(r"\(.*", "", nice)
        cols = ["%s=%d" % (w, c[w]) for w in WATCH if c[w]]
        print("%-46s total=%-6d %s" % (nice[:46], c["_total"], " ".join(cols)))
        tot.update(c)
    print("%-46s total=%-6d %s" % ("ALL KERNELS", tot["_total"], " ".join("%s=%d" % (w, tot[w]) for w in WATCH)))


if __name__ == "__main__":
    main()
