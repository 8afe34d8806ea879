"""Developer diagnostic (not a test): exercises every kernel once on the GPU and prints how it
compares with the oracles, without stopping at the first mismatch."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import supersdr_b200 as S
from oracle import c_oracle, tier_p, tier_u

S.init()
rng = np.random.default_rng(1)

print("== tables / plan")
for N in (256, 512, 1024, 2048, 4096, 8192, 16384):
    wb = S.WaterfallBank(N, 1, 1)
    tw, th, plan = wb.tables()
    ok_tw = np.array_equal(tw.view(np.float32), c_oracle.twiddle_table(N).view(np.float32))
    ok_th = np.array_equal(th, c_oracle.thresholds(N, -10.0))
    print(N, plan, c_oracle.fft_plan(N), "tw", ok_tw, "thr", ok_th)
    wb.close()

print("== waterfall vs C oracle")
for N, B, n in ((256, 33, 3), (512, 9, 2), (1024, 5, 1), (1024, 8, 10), (2048, 5, 4), (4096, 3, 2), (8192, 3, 2), (16384, 3, 2), (16384, 150, 1)):
    iq = tier_u.synth_batch(B, n, N, seed=N + B)
    wb = S.WaterfallBank(N, B, n)
    wb.set_display(zoom=2)
    t = time.time(); res = wb.process(iq); dt = time.time() - t
    ref = c_oracle.wf_rows(iq, zoom=2, threads=8)
    spec_ok = np.array_equal(res["spectrum"], ref["spectrum"])
    px_ok = np.array_equal(res["pixels"], ref["pixels"])
    col_ok = np.array_equal(res["colour"], ref["colour"])
    sc = np.stack([res["scalars"][k] for k in ("low_clip_db", "high_clip_db", "dynamic_range", "wf_min_db", "wf_max_db")], 1)
    sc_ok = np.array_equal(sc, ref["scalars"])
    print("N=%d B=%d n=%d spectrum %s colour %s pixels %s scalars %s  (%.1f ms)" % (N, B, n, spec_ok, col_ok, px_ok, sc_ok, dt * 1e3))
    if not spec_ok:
        d = res["spectrum"] != ref["spectrum"]
        print("   spectrum mismatches:", d.sum(), "of", d.size, "first:", np.argwhere(d)[:5].tolist(),
              res["spectrum"][d][:5], ref["spectrum"][d][:5])
    if spec_ok and not sc_ok:
        print("   scalars", sc[:2], ref["scalars"][:2])
    # s16be path
    if N in (1024, 16384) and n <= 2:
        q = np.clip(np.rint(np.stack([iq.real, iq.imag], -1)), -32768, 32767).astype(">i2")
        wire = q.view(np.uint8).reshape(B, n, N, 4)
        iq_q = (q[..., 0].astype(np.float32) + 1j * q[..., 1].astype(np.float32)).astype(np.complex64)
        r2 = wb.process(wire); ref2 = c_oracle.wf_rows(iq_q, zoom=2, threads=8)
        print("   s16be pixels", np.array_equal(r2["pixels"], ref2["pixels"]), "spectrum", np.array_equal(r2["spectrum"], ref2["spectrum"]))
    wb.close()

print("== colorrow (Tier P) vs numpy restatement")
for W, B, n in ((1024, 7, 1), (1024, 16, 10), (256, 40, 3), (16384, 3, 100), (2048, 4, 7)):
    lines = np.clip(rng.normal(120, 8, (B, n, W)), 0, 255).astype(np.uint8)
    wb = S.WaterfallBank(W, B, n); wb.set_display(zoom=5, delta_low_db=-3, delta_high_db=4)
    res = wb.colorrow(lines)
    ok = True
    for b in range(B):
        st = tier_p.ColourState(); st.zoom = 5; st.delta_low_db = -3; st.delta_high_db = 4
        spec, col, px = tier_p.waterfall_line(lines[b], st)
        ok &= np.array_equal(spec, res["spectrum"][b]) and np.array_equal(col, res["colour"][b]) and np.array_equal(px, res["pixels"][b])
        ok &= np.float32(st.low_clip_db) == res["scalars"]["low_clip_db"][b] and np.float32(st.wf_max_db) == res["scalars"]["wf_max_db"][b]
    print("W=%d B=%d n=%d ->" % (W, B, n), ok)
    wb.close()

print("== demod vs float64 oracle")
for mode in ("usb", "lsb", "cw", "am", "nbfm"):
    for hang, on in ((False, True), (True, True), (False, False)):
        B, n = 6, 512 * 12
        p = dict(mode=mode, hang=hang, on=on, decay=1000 if mode == "cw" else 4000, slope=6 if hang else 0)
        bank = S.DemodBank(B, n)
        bank.set_all(**p)
        iq = np.stack([tier_u.synth_demod_iq(mode, n, seed=10 + b, level=0.1 / (b + 1)) for b in range(B)])
        # stream in two calls
        r1 = bank.process(iq[:, : n // 2].copy()); r2 = bank.process(iq[:, n // 2:].copy())
        got = np.concatenate([r1["pcm_f32"], r2["pcm_f32"]], 1); rssi = np.concatenate([r1["rssi"], r2["rssi"]], 1)
        gi = np.concatenate([r1["pcm_i16"], r2["pcm_i16"]], 1)
        worst = 0; worst_r = 0; li = 0
        for b in range(B):
            st = tier_u.DemodState(); ref, rr = tier_u.demod(iq[b], tier_u.DemodParams(**p), st)
            e = np.sqrt(np.mean((got[b] - ref) ** 2)) / max(np.sqrt(np.mean(ref ** 2)), 1e-30)
            worst = max(worst, e); worst_r = max(worst_r, np.abs(rssi[b] - rr).max())
            li = max(li, np.abs(gi[b].astype(int) - tier_u.pcm_to_i16(ref).astype(int)).max())
        print("%-5s hang=%d on=%d  rel-RMS err %.2e  rssi err %.2e dB  int16 max diff %d" % (mode, hang, on, worst, worst_r, li))
        bank.close()

print("== interp vs tier_p.play_buffer")
B, n = 5, 512
ib = S.InterpBank(B, 4, max_samples=n)
sts = [tier_p.InterpState() for _ in range(B)]
ok = True; md = 0
for it in range(6):
    x = rng.integers(-32768, 32768, (B, n)).astype(np.int16)
    vol = rng.integers(0, 16, B) * 10.0; bal = rng.uniform(-1, 1, B).astype(np.float32)
    out, mono = ib.process(x, vol, bal, want_mono=True)
    for b in range(B):
        buf, o2 = tier_p.play_buffer(x[b], sts[b], float(vol[b]), float(bal[b]))
        md = max(md, np.abs(out[b].astype(int) - o2.astype(int)).max())
        ok &= np.allclose(mono[b], buf, rtol=0, atol=1e-7)
print("interp max int16 diff", md, "mono close", ok)
f = S.filtering(6000, 48000); x = rng.standard_normal(5000)
print("filtering.lowpass max err", np.abs(f.lowpass(x) - np.convolve(x, f.h, "valid")).max())
raw = rng.integers(-32768, 32768, 2000).astype(">i2")
print("unpack ok", np.array_equal(S.unpack_iq(raw.tobytes()), (raw[0::2].astype(np.float32) + 1j * raw[1::2].astype(np.float32)).astype(np.complex64)))

print("== quick timing, config 2 (4096 x 10 x 16384 complex64 resident)")
B, n, N = 4096, 10, 16384
buf = S.DeviceBuffer(B * n * N * 8); px = S.DeviceBuffer(B * N)
S._lib.check(S.lib.ssdr_synth_iq_dev(buf.ptr, 0, B, n, N, 1234))
wb = S.WaterfallBank(N, B, n)
for _ in range(3): wb.time_dev(buf.ptr, 0, px.ptr, 1)
ms = wb.time_dev(buf.ptr, 0, px.ptr, 5) / 5
gb = (B * n * N * 8 + B * N) / 1e9
print("wf kernel %.3f ms  %.1f Gsamples/s  %.0f GB/s (%.1f%% of 6545)" % (ms, B * n * N / ms / 1e6, gb / ms * 1e3, gb / ms * 1e3 / 65.45))
pix = px.download(np.uint8, (B, N)); print("pixels hist head", np.bincount(pix.ravel(), minlength=256)[:8], "max", pix.max())
# check 3 channels of the big run against the oracle on the same bits
for ch in (0, 1777, 4095):
    x = buf.download(np.complex64, (1, n, N), offset_bytes=ch * n * N * 8)
    ref = c_oracle.wf_rows(x)
    print("   ch", ch, "pixels equal oracle:", np.array_equal(ref["pixels"][0], pix[ch]))
wb.close(); buf.free(); px.free()
print("== quick timing, demod config 3 (4096 ch x 64 frames)")
B, n = 4096, 512 * 64
buf = S.DeviceBuffer(B * n * 8); out = S.DeviceBuffer(B * n * 4)
S._lib.check(S.lib.ssdr_synth_iq_dev(buf.ptr, 0, B, 1, n, 99))
db = S.DemodBank(B, n); db.set_all(mode="usb", lc=300, hc=2700)
for _ in range(2): db.time_dev(buf.ptr, 0, n, out.ptr, None, 1)
ms = db.time_dev(buf.ptr, 0, n, out.ptr, None, 3) / 3
print("demod kernel %.3f ms  %.1f Gsamples/s  %.0f GB/s" % (ms, B * n / ms / 1e6, B * n * 12 / ms / 1e6))
print("launches", S.lib.ssdr_launch_count())
