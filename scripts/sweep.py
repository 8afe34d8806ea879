#!/usr/bin/env python
"""BASELINE config 5: FFT size sweep x batch sweep on one GPU (device-resident IQ, n_avg = 1 unless --n-avg):
achieved algorithmic HBM GB/s (B*n*N*8 in + B*N out) vs the measured copy peak.  Prints one JSON line per point.
    python scripts/sweep.py [--sizes 256,...,16384] [--batches 1,64,4096,65536] [--n-avg 1] [--fmt cf32|s16be]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import supersdr_b200 as S


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="256,512,1024,2048,4096,8192,16384,32768,65536")
    ap.add_argument("--batches", default="1,16,256,4096,65536")
    ap.add_argument("--n-avg", type=int, default=1)
    ap.add_argument("--fmt", default="cf32", choices=["cf32", "s16be"])
    ap.add_argument("--max-bytes", type=float, default=8e9)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--gpus", type=int, default=1, help="N > 1: one child process per GPU runs the same sweep on its own shard of "
                    "`batch` channels per GPU (channels are independent: no collective); prints the per-point aggregate")
    a = ap.parse_args()
    if a.gpus > 1:
        import subprocess
        cmd = [sys.executable, os.path.abspath(__file__), "--sizes", a.sizes, "--batches", a.batches, "--n-avg", str(a.n_avg),
               "--fmt", a.fmt, "--max-bytes", str(a.max_bytes), "--iters", str(a.iters)]
        procs = [subprocess.Popen(cmd + ["--device", str(d)], stdout=subprocess.PIPE, text=True) for d in range(a.gpus)]
        outs = [[json.loads(l) for l in p.communicate()[0].splitlines() if l.startswith("{")] for p in procs]
        for rows in zip(*outs):
            r0 = dict(rows[0])
            r0.update({"n_gpus": a.gpus, "batch_per_gpu": r0["batch"], "ms": max(r["ms"] for r in rows),
                       "msamples_per_s": round(sum(r["msamples_per_s"] for r in rows), 1), "gbs": round(sum(r["gbs"] for r in rows), 1),
                       "frac_of_measured_hbm": round(min(r["frac_of_measured_hbm"] for r in rows), 4),
                       "frac_per_gpu": [r["frac_of_measured_hbm"] for r in rows]})
            print(json.dumps(r0), flush=True)
        return
    S.init(a.device)
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    fmt = S.SSDR_IQ_CF32 if a.fmt == "cf32" else S.SSDR_IQ_S16BE
    sb = 8 if a.fmt == "cf32" else 4
    for N in [int(x) for x in a.sizes.split(",")]:
        for B in [int(x) for x in a.batches.split(",")]:
            in_bytes = B * a.n_avg * N * sb
            if in_bytes > a.max_bytes:
                continue
            iq = S.DeviceBuffer(in_bytes)
            px = S.DeviceBuffer(B * N)
            S._lib.check(S.lib.ssdr_synth_iq_dev(iq.ptr, fmt, B, a.n_avg, N, 77))
            bank = S.WaterfallBank(N, B, a.n_avg)
            for _ in range(3):
                bank.time_dev(iq.ptr, fmt, px.ptr, 1)
            # inputs smaller than L2 (126 MB) are cache-resident between iterations: flagged in the output
            ms = bank.time_dev(iq.ptr, fmt, px.ptr, a.iters) / a.iters
            alg = in_bytes + B * N
            print(json.dumps({"nfft": N, "batch": B, "n_avg": a.n_avg, "fmt": a.fmt, "ms": round(ms, 5),
                              "msamples_per_s": round(B * a.n_avg * N / ms / 1e3, 1), "gbs": round(alg / ms / 1e6, 1),
                              "frac_of_measured_hbm": round(alg / ms / 1e6 / peak, 4),
                              "l2_resident": in_bytes < 126e6,
                              # N > 16384: three-kernel path (default): front pass + scratch round trip through HBM, 3x the input bytes
                              # move; fused kernel (SSDR_WF_BIG=fused): the scratch stays in L2, but slower
                              "big_path": ("fused" if os.environ.get("SSDR_WF_BIG") == "fused" else "3k") if N > 16384 else None,
                              "hbm_passes": (1 if os.environ.get("SSDR_WF_BIG") == "fused" else 3) if N > 16384 else 1}), flush=True)
            bank.close(); iq.free(); px.free()


if __name__ == "__main__":
    main()
