#!/usr/bin/env python
"""Tier-P entry (finished uint8 W/F lines in, utils_supersdr.py:783-813,881-888) on device-resident lines: achieved HBM
GB/s of wf_colorrow_kernel (bytes = B*n*W in + B*W pixels out).  Prints one JSON line per shape.
    python scripts/colorrow_bw.py [--shapes 16384x4096x10,1024x65536x10,...]"""
import argparse, ctypes, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import supersdr_b200 as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="16384x4096x10,8192x8192x10,1024x65536x10,1024x65536x1,16384x4096x1")
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    S.init(0)
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    for shp in a.shapes.split(","):
        W, B, n = (int(x) for x in shp.split("x"))
        lines = S.DeviceBuffer(B * n * W)
        px = S.DeviceBuffer(B * W)
        # random bytes over the WHOLE buffer (a 16 MB block tiled): constant rows would take the row stage's shortcut.  (Until
        # round 2's last session only the first 16 MB were random -- the numbers of that time flatter the large shapes.)
        rng = np.random.default_rng(1)
        host = rng.integers(60, 200, min(B * n * W, 1 << 24)).astype(np.uint8)
        for o in range(0, B * n * W, host.size):
            S._lib.check(S.lib.ssdr_memcpy_h2d(ctypes.c_void_p(lines.ptr.value + o), S._lib.ptr(host), min(host.size, B * n * W - o)))
        bank = S.WaterfallBank(W, B, n)
        call = lambda: S._lib.check(S.lib.ssdr_wf_colorrow_u8_dev(bank._h, lines.ptr, px.ptr, None, None, None))
        for _ in range(3):
            call()
        bank.sync()
        t0 = time.perf_counter()
        for _ in range(a.iters):
            call()
        bank.sync()
        ms = (time.perf_counter() - t0) / a.iters * 1e3
        byts = B * n * W + B * W
        print(json.dumps({"W": W, "batch": B, "n_avg": n, "ms": round(ms, 4), "gbs": round(byts / ms / 1e6, 1),
                          "frac_of_measured_hbm": round(byts / ms / 1e6 / peak, 4), "mlines_per_s": round(B * n / ms / 1e3, 1),
                          "l2_resident": B * n * W < 126e6}), flush=True)
        bank.close(); lines.free(); px.free()


if __name__ == "__main__":
    main()
