#!/usr/bin/env python
"""Register-file READ pressure of a profiled kernel (developer tool): from an .ncu-rep's source page, weight every SASS
instruction's vector-register source operands (not served by the .reuse operand cache; a packed .F32x2 / .64 operand
counts two registers, a .128 store four) by its executed count.  Packed fp32 on sm_100 is bound by operand delivery,
about one 32-bit register per lane per cycle per sub-partition (profiles/r2c_ubench_ffma2_modifiers.txt), so
reads / (4 sub-partitions) is the cycle floor per SM.

    python scripts/rf_reads.py report.ncu-rep [samples_per_launch]"""
import collections
import csv
import io
import re
import subprocess
import sys


def src_regs(text):
    """(#source registers read from the RF, opcode) of one SASS line."""
    parts = text.strip().rstrip(";").split(None, 1)
    if parts and parts[0].startswith("@"):
        pred, rest = parts[0], parts[1] if len(parts) > 1 else ""
        parts = rest.split(None, 1)
    if not parts:
        return 0, ""
    op = parts[0]
    ops = parts[1] if len(parts) > 1 else ""
    base = op.split(".")[0]
    toks = [t.strip() for t in re.split(r",(?![^\[]*\])", ops)]
    store = base in ("STS", "STG", "STL", "ST", "RED", "ATOMS", "ATOMG", "ATOM", "STTM", "UTCBAR")
    branch = base in ("BRA", "BSSY", "BSYNC", "EXIT", "BAR", "NOP", "CALL", "RET", "WARPSYNC", "SYNCS", "DEPBAR", "MEMBAR", "ERRBAR")
    srcs = toks if (store or branch) else toks[1:]
    width = 4 if ".128" in op else 2 if ".64" in op else 1
    n = 0
    for i, t in enumerate(srcs):
        for m in re.finditer(r"(?<![A-Z])R(\d+)((?:\.[A-Za-z0-9_]+)*)", t):
            mods = m.group(2)
            if ".reuse" in mods:
                continue
            w = 1
            if "F32x2" in mods or ".64" in mods:
                w = 2
            if store and i == len(srcs) - 1 and "[" not in t:
                w = width                     # store data
            if base in ("STTM",) and "[" not in t:
                w = int(re.search(r"x(\d+)", op).group(1)) if re.search(r"x(\d+)", op) else 1
            n += w
    return n, base


def main():
    rep = sys.argv[1]
    nsamples = float(sys.argv[2]) if len(sys.argv) > 2 else None
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr = rows[hi]
    data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
    reads, instr = collections.Counter(), collections.Counter()
    for r in data:
        n, base = src_regs(r[iS])
        e = int(r[iE])
        reads[base] += n * e
        instr[base] += e
    tot_r, tot_i = sum(reads.values()), sum(instr.values())
    print("warp-instructions %d, register source reads %d (%.2f per instruction)" % (tot_i, tot_r, tot_r / max(tot_i, 1)))
    if nsamples:
        print("per sample: %.2f lane-instructions, %.2f register reads" % (tot_i * 32 / nsamples, tot_r * 32 / nsamples))
    for op, c in reads.most_common(16):
        line = "  %-8s %5.1f%% of reads  %5.1f%% of instr  %.2f reads/instr" % (op, 100 * c / tot_r, 100 * instr[op] / tot_i, c / max(instr[op], 1))
        if nsamples:
            line += "  %6.2f reads/sample" % (c * 32 / nsamples)
        print(line)


if __name__ == "__main__":
    main()
