#!/usr/bin/env python
"""Developer tool: static opcode mix of the longest loop (the frame loop) of a kernel in an object file.
usage: sass_mix.py <obj> <mangled-function-substring>"""
import re, subprocess, collections, sys
obj, pat = sys.argv[1], sys.argv[2]
names = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funs = sorted(set(re.findall(r"Function : (\S+)", names)))
fun = [f for f in funs if pat in f][0]
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
addr = {a: i for i, (a, _) in enumerate(ins)}
loops = []
for i, (a, t) in enumerate(ins):
    m = re.search(r"BRA\S*\s+(?:\S+,\s*)?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr: loops.append((addr[tgt], i))
# frame loop = the longest loop that is not the outermost (channel) loop
loops.sort(key=lambda x: x[1] - x[0], reverse=True)
lo, hi = loops[1] if len(loops) > 1 and loops[0][0] <= loops[1][0] and loops[1][1] <= loops[0][1] else loops[0]
body = ins[lo:hi + 1]
ops = collections.Counter()
for _, t in body:
    p = t.split()
    op = p[1] if p[0].startswith("@") else p[0]
    ops[op.split(".")[0]] += 1
print("%s: %d instructions total, frame loop [%d,%d] = %d" % (fun[:60], len(ins), lo, hi, len(body)))
print("  " + "  ".join("%s:%d" % kv for kv in ops.most_common(30)))
fma = sum(ops[k] for k in ("FADD2", "FMUL2", "FFMA2"))
print("  packed fp32: %d   scalar fp32: %d   other: %d" % (fma, sum(ops[k] for k in ("FADD", "FMUL", "FFMA")), len(body) - fma - sum(ops[k] for k in ("FADD", "FMUL", "FFMA"))))
