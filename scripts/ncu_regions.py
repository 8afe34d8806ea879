#!/usr/bin/env python
"""Per-region stall breakdown from the ncu source page (regions = blocks of K SASS instructions, or split at BAR)."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; K = int(sys.argv[2]) if len(sys.argv) > 2 else 300
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
iS, iE, iN = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot_s = sum(int(r[iN]) for r in data) or 1
tot_e = sum(int(r[iE]) for r in data) or 1
# split regions at barriers
regions = []; cur = []
for r in data:
    cur.append(r)
    op = r[iS].split()
    if any(o.startswith("BAR") or o.startswith("WARPSYNC") for o in op[:2]) or len(cur) >= K:
        regions.append(cur); cur = []
if cur: regions.append(cur)
idx = 0
for reg in regions:
    e = sum(int(r[iE]) for r in reg); s = sum(int(r[iN]) for r in reg)
    if s * 100 > tot_s or e * 100 > tot_e:
        st = collections.Counter()
        for r in reg:
            for i, h in stall_cols:
                try: st[h[6:]] += int(r[i])
                except ValueError: pass
        top = ", ".join("%s %.0f%%" % (k, 100 * v / max(s, 1)) for k, v in st.most_common(5))
        print("@%5d +%4d  instr %5.1f%%  samples %5.1f%%  | %s | ends: %s" % (idx, len(reg), 100 * e / tot_e, 100 * s / tot_s, top, reg[-1][iS].strip()[:40]))
    idx += len(reg)
