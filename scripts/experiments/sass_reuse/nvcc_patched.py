#!/usr/bin/env python
"""nvcc -c with one extra step between ptxas and fatbinary: tools/sass_reuse.py on the cubin.

nvcc has no hook after ptxas, so this asks nvcc for its own command list (`-dryrun --keep`), runs those commands
unchanged and inserts the peephole right after the ptxas line.  Everything else (front end, cicc, ptxas flags,
fatbinary, host compile) is exactly what `nvcc -c` would do with the same flags.

usage: nvcc_patched.py --keep-dir DIR --patch PATTERN[,PATTERN...] [--max-dist N] -- <nvcc arguments of a -c compile>"""
import os
import re
import subprocess
import sys


def main():
    a = sys.argv[1:]
    cut = a.index("--")
    own, nv = a[:cut], a[cut + 1:]
    keep = own[own.index("--keep-dir") + 1]
    pats = own[own.index("--patch") + 1].split(",")
    extra = ["--max-dist", own[own.index("--max-dist") + 1]] if "--max-dist" in own else []
    os.makedirs(keep, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    dry = subprocess.run([nvcc, "-dryrun", "--keep", "--keep-dir", keep] + nv, capture_output=True, text=True, check=True).stderr
    tool = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sass_reuse.py")
    script = ["set -e"]
    patched = False
    for l in dry.splitlines():
        if not l.startswith("#$ "):
            continue
        l = l[3:]
        if re.match(r"^[A-Za-z_]+=", l) and not l.startswith(("PATH=", "LD_LIBRARY_PATH=", "CICC_PATH=", "NVVMIR_LIBRARY_DIR=")):
            continue                       # nvcc's internal variables that no command line below refers to
        script.append(l)
        m = re.match(r'^ptxas .*-o "([^"]+\.cubin)"', l)
        if m:
            cubin = m.group(1)
            script.append('%s %s "%s" "%s.patched" %s %s --report' % (sys.executable, tool, cubin, cubin, " ".join(pats), " ".join(extra)))
            script.append('mv "%s" "%s.ptxas" && mv "%s.patched" "%s"' % (cubin, cubin, cubin, cubin))
            patched = True
    assert patched, "no ptxas step in nvcc's command list"
    path = os.path.join(keep, "build.sh")
    open(path, "w").write("\n".join(script) + "\n")
    subprocess.run(["bash", path], check=True)


if __name__ == "__main__":
    main()
