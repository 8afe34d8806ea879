"""Small invocations of every kernel, for compute-sanitizer (memcheck / racecheck / synccheck).
usage: compute-sanitizer --tool <tool> python scripts/gpu_sanitize.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import supersdr_b200 as S
from oracle import tier_u

S.init(0)
rng = np.random.default_rng(0)
for N, B, n in ((256, 9, 2), (1024, 3, 2), (4096, 3, 2), (8192, 2, 2), (16384, 2, 3), (32768, 1, 2), (65536, 1, 1)):
    iq = tier_u.synth_batch(B, n, N, seed=N)
    bank = S.WaterfallBank(N, B, n)
    r = bank.process(iq)
    assert r["pixels"].shape == (B, N)
    if N <= 16384:
        lines = rng.integers(0, 256, (B, n, N)).astype(np.uint8)
        bank.colorrow(lines)
    bank.close()
    print("waterfall", N, "ok", flush=True)
modes = ("am", "lsb", "usb", "cw", "nbfm", "usb", "usb")      # 7 channels: partial quads, three filters
x = np.stack([tier_u.synth_demod_iq(m, 512 * 4, seed=1) for m in modes])
for eng in ("ffma", "tcgen05"):
    dm = S.DemodBank(len(modes), 512 * 4, engine=eng)
    dm.set_params(0, [S.demod_params(m) for m in modes])
    dm.process(x); dm.process(x)
    dm.close()
print("demod ok", flush=True)
ib = S.InterpBank(3, 4, max_samples=512)
ib.process(rng.integers(-20000, 20000, (3, 512)).astype(np.int16), volume=80, balance=0.2)
ib.close()
print("interp ok", flush=True)
img = S.WaterfallImage(2, 8, 1024)
for _ in range(6):
    img.push(rng.uniform(0, 254, (2, 1024)).astype(np.float32))
img.image(True, True); img.trace(15, 100); img.set_white_flag()
img.close()
rs = S.ResampleLine(64, 27)
rs.process(rng.integers(-20000, 20000, (2, 512)).astype(np.int16), 90, -0.1)
print("image / resample ok", flush=True)
