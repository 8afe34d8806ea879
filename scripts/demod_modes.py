#!/usr/bin/env python
"""Demodulator throughput per mode and FIR engine on one GPU (device-resident synthetic IQ): uniform batches of each
mode, so the cost classes behind the tcgen05 engine's round order (capi.cu: demod_plan_rounds) can be read off.
Prints one JSON line per (mode, engine).
    python scripts/demod_modes.py [--batch 4096] [--frames 32] [--modes am,lsb,usb,cw,nbfm] [--hang 0|1]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import supersdr_b200 as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--modes", default="am,lsb,usb,cw,nbfm")
    ap.add_argument("--hang", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    S.init(0)
    B, n = a.batch, 512 * a.frames
    iq = S.DeviceBuffer(B * n * 8)
    out = S.DeviceBuffer(B * n * 4)
    S._lib.check(S.lib.ssdr_synth_iq_dev(iq.ptr, S.SSDR_IQ_CF32, B, 1, n, 99))
    for mode in a.modes.split(","):
        bank = S.DemodBank(B, n)
        bank.set_params(0, [S.demod_params(mode, hang=bool(a.hang))] * B)
        for eng in ("ffma", "tcgen05"):
            bank.set_engine(eng)
            for _ in range(3):
                bank.time_dev(iq.ptr, S.SSDR_IQ_CF32, n, out.ptr, None, 1)
            ms = bank.time_dev(iq.ptr, S.SSDR_IQ_CF32, n, out.ptr, None, a.iters) / a.iters
            print(json.dumps({"mode": mode, "engine": eng, "batch": B, "frames": a.frames, "hang": a.hang, "ms": round(ms, 4),
                              "gsamples_per_s": round(B * n / ms / 1e6, 1)}), flush=True)
        bank.close()
    iq.free(); out.free()


if __name__ == "__main__":
    main()
