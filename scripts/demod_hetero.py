#!/usr/bin/env python
"""Demodulator throughput against the number of channels that share a filter (device-resident synthetic IQ): a bank of
`--batch` USB channels whose pass-band widths take G distinct values (hc = 2700 + g Hz, g = ch % G), i.e. batch / G
channels per filter.  The tcgen05 engine shares one Toeplitz operand per CTA round (four tiles of four channels), so
few channels per filter leave tiles and warps empty; the FFMA engine does not care.  Prints one JSON line per
(G, engine) with the plan's tile fill: the data behind the AUTO rule (capi.cu: demod_launch_block).
    python scripts/demod_hetero.py [--batch 4096] [--frames 32] [--per-filter 1,2,4,8,16,64]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import supersdr_b200 as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--per-filter", default="1,2,4,8,16,64")
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    S.init(0)
    B, n = a.batch, 512 * a.frames
    iq = S.DeviceBuffer(B * n * 8)
    out = S.DeviceBuffer(B * n * 4)
    S._lib.check(S.lib.ssdr_synth_iq_dev(iq.ptr, S.SSDR_IQ_CF32, B, 1, n, 99))
    for per in [int(x) for x in a.per_filter.split(",")]:
        G = max(1, B // per)
        uniq = [S.demod_params("usb", lc=300, hc=2700 + g) for g in range(G)]
        params = [uniq[ch % G] for ch in range(B)]
        bank = S.DemodBank(B, n)
        bank.set_params(0, params)
        fill = S.demod_plan(params)[3]
        for eng in ("ffma", "tcgen05", "auto"):
            bank.set_engine(eng)
            for _ in range(3):
                bank.time_dev(iq.ptr, S.SSDR_IQ_CF32, n, out.ptr, None, 1)
            ms = bank.time_dev(iq.ptr, S.SSDR_IQ_CF32, n, out.ptr, None, a.iters) / a.iters
            print(json.dumps({"channels_per_filter": per, "filters": G, "engine": eng, "tile_fill": round(float(fill), 3), "batch": B,
                              "frames": a.frames, "ms": round(ms, 4), "gsamples_per_s": round(B * n / ms / 1e6, 1)}), flush=True)
        bank.close()
    iq.free(); out.free()


if __name__ == "__main__":
    main()
