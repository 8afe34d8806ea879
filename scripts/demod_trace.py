#!/usr/bin/env python
"""Developer tool: phase timeline of demod_tc_kernel (CTA 0, every warp) from a -DSSDR_TRACE build, BASELINE config 3 shape.
    SSDR_B200_LIB=build/exp/libssdr_dtrace.so python scripts/demod_trace.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import supersdr_b200 as S

S.init(0)
B, n = 4096, 512 * 64
iq = S.DeviceBuffer(B * n * 8); out = S.DeviceBuffer(B * n * 4)
S._lib.check(S.lib.ssdr_synth_iq_dev(iq.ptr, S.SSDR_IQ_CF32, B, 1, n, 99))
bank = S.DemodBank(B, n)
bank.set_params(0, [S.demod_params("usb", 300, 2700)] * B)
bank.set_engine("tcgen05")
for _ in range(2):
    ms = bank.time_dev(iq.ptr, S.SSDR_IQ_CF32, n, out.ptr, None, 1)
print("ms", ms, "Gsamples/s", B * n / ms / 1e6)
F, W, P = 96, 16, 8
buf = (C.c_longlong * (F * W * P))()
S.lib.ssdr_debug_demod_trace(buf)
t = np.frombuffer(buf, dtype=np.int64).reshape(F, W, P).astype(np.float64)
names = ["fences + arrive (+ MMA issue)", "back end of frame b-1", "wait for the MMAs", "history rows", "IQ loads + mixer", "operand stores"]
fs = slice(8, 56)
d = np.diff(t[fs, :, :7], axis=2)
print("frame period per warp (cycles):", np.diff(t[fs, :, 0], axis=0).mean())
for k, nm in enumerate(names):
    print("%-32s %7.0f   per tile %s" % (nm, d[:, :, k].mean(), "  ".join("%6.0f" % d[:, 4 * i:4 * i + 4, k].mean() for i in range(4))))
if os.environ.get("TRACE_RAW"):
    for f in (20, 21, 22):
        b0 = t[f, :, 0].min()
        print("frame %d (rel. cycles): per warp T0 (rows ready), T1 (arrived / issued), T2 (back end done), T3 (MMAs done), T6 (rows of next frame stored)" % f)
        for k in (0, 1, 2, 3, 6):
            print("  T%d %s" % (k, " ".join("%6.0f" % (v - b0) for v in t[f, :, k])))
