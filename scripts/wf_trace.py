#!/usr/bin/env python
"""Developer tool: phase timeline of wf_fft_kernel<14> (CTA 0, every warp) from a -DSSDR_TRACE build.
    SSDR_B200_LIB=build/exp/libssdr_trace.so python scripts/wf_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import supersdr_b200 as S

S.init(0)
N, B, n = 16384, 4096, 10
iq = S.DeviceBuffer(B * n * N * 8); px = S.DeviceBuffer(B * N)
S._lib.check(S.lib.ssdr_synth_iq_dev(iq.ptr, S.SSDR_IQ_CF32, B, n, N, 77))
bank = S.WaterfallBank(N, B, n)
for _ in range(2):
    ms = bank.time_dev(iq.ptr, S.SSDR_IQ_CF32, px.ptr, 1)
print("ms", ms)
F, W, P = 40, 16, 24
buf = (C.c_longlong * (F * W * P))(); fr = C.c_int()
S.lib.ssdr_debug_wf_trace(buf, C.byref(fr))
full = np.frombuffer(buf, dtype=np.int64).reshape(F, W, P).astype(np.float64)
t = full[:, :, :10]
names = ["math b0", "buffer_free wait", "store b0 + issue loads", "math b1 (incl. load wait)", "store b1", "group_sync wait", "stagger spin",
         "pass_mid", "pass_last + quantiser"]
fs = slice(12, 38)
print("frame period (cycles):", np.diff(t[fs, 0, 0]).mean())
d = np.diff(t[fs], axis=2)                        # [frames][warps][9]
print("%-28s %8s   per stagger level 0..3" % ("phase", "mean"))
for k, nm in enumerate(names):
    lv = [d[:, l * 4:(l + 1) * 4, k].mean() for l in range(4)]
    print("%-28s %8.0f   %s" % (nm, d[:, :, k].mean(), "  ".join("%6.0f" % v for v in lv)))
gap = t[fs.start + 1:fs.stop + 1, :, 0] - t[fs, :, 9]
print("%-28s %8.0f" % ("end of frame -> next T0", gap.mean()))
# offsets of each point relative to the earliest T0 of the frame, per level
base = t[fs, :, 0].min(axis=1)[:, None, None]
rel = t[fs] - base
print("offsets from the frame's first T0, by level:")
for k in range(10):
    print("  T%d  %s" % (k, "  ".join("%6.0f" % rel[:, l * 4:(l + 1) * 4, k].mean() for l in range(4))))
print("gap T9 -> next T0 by frame index within the channel (9 = colour stage follows):")
for r in range(10):
    idx = [f for f in range(10, 38) if f % 10 == r]
    g = np.array([t[f + 1, :, 0] - t[f, :, 9] for f in idx])
    print("  f=%d  %7.0f" % (r, g.mean()))
print("block T0->T3 (b0: load wait, math, buffer wait, store) per warp:", np.round((t[fs, :, 3] - t[fs, :, 0]).mean(axis=0)))
print("block T5->T7 (sync + stagger) per warp:", np.round((t[fs, :, 7] - t[fs, :, 5]).mean(axis=0)))
print("arrival T5 per warp (rel):", np.round(rel[:, :, 5].mean(axis=0)))
print("release T7 per warp (rel):", np.round(rel[:, :, 7].mean(axis=0)))

# inside pass_last: T8 -> loads issued/hook (10) -> dft32 done (11) -> quantiser groups (12, 13, 14) -> T9
q = full[12:38]
seq = [8, 10, 11, 12, 13, 14, 9]
lab = ["loads + hook", "dft32", "quantiser group 0", "group 1", "group 2", "group 3"]
print("inside pass_last, by level:")
for a, b, nm in zip(seq[:-1], seq[1:], lab):
    dd = q[:, :, b] - q[:, :, a]
    print("  %-20s %s" % (nm, "  ".join("%6.0f" % dd[:, l * 4:(l + 1) * 4].mean() for l in range(4))))

# row stage (once per channel): its time stamps sit in slots 10..15 of the channel's NEXT frame record (f % 10 == 0)
rows = [f for f in range(10, 38) if f % 10 == 0]
if rows:
    c = np.array([full[f, :, 16:22] for f in rows])            # [channels][warps][6]
    prev9 = np.array([full[f - 1, :, 9] for f in rows])
    nxt0 = np.array([full[f, :, 0] for f in rows])
    lab = ["entry (accumulator reload)", "max", "histogram (32 shared atomics per thread)", "rank scan, min, lerp", "key -> colour table, row staged", "row stores (32 per thread)"]
    print("row stage, cycles (mean over warps):")
    print("  %-44s %7.0f" % ("last quantiser -> entry", (c[:, :, 0] - prev9).mean()))
    for k in range(5):
        print("  %-44s %7.0f" % (lab[k + 1], (c[:, :, k + 1] - c[:, :, k]).mean()))
    print("  %-44s %7.0f" % ("end -> first wait of the next channel (T0)", (nxt0 - c[:, :, 5]).mean()))
