"""GPU parity tests (through the C ABI) for the waterfall path: bit-exact against the oracle."""
import os

import numpy as np
import pytest

from oracle import c_oracle, tier_p, tier_u

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
SC = ("low_clip_db", "high_clip_db", "dynamic_range", "wf_min_db", "wf_max_db")


def _sc(res):
    return np.stack([res["scalars"][k] for k in SC], 1)


def test_spec_tables_equal_oracle(ssdr):
    for N in (256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        b = ssdr.WaterfallBank(N, 1, 1)
        tw, thr, plan = b.tables()
        assert plan == c_oracle.fft_plan(N)
        assert np.array_equal(tw.view(np.float32), c_oracle.twiddle_table(N).view(np.float32))
        assert np.array_equal(thr, c_oracle.thresholds(N, -10.0))
        assert np.array_equal(b.window_table(), c_oracle.window_table(N))
        b.close()


@pytest.mark.parametrize("N,B,n", [(256, 33, 3), (512, 9, 2), (1024, 1, 1), (1024, 8, 10), (2048, 5, 4),
                                   (4096, 3, 2), (8192, 3, 2), (16384, 3, 2), (16384, 149, 1),
                                   (32768, 3, 2), (32768, 150, 1), (65536, 2, 3), (65536, 75, 1)])
def test_waterfall_bit_exact_vs_c_oracle(ssdr, N, B, n):
    iq = tier_u.synth_batch(B, n, N, seed=N + B)
    bank = ssdr.WaterfallBank(N, B, n)
    bank.set_display(zoom=2)
    res = bank.process(iq)
    ref = c_oracle.wf_rows(iq, zoom=2, threads=8)
    assert np.array_equal(res["spectrum"], ref["spectrum"])      # integer byte sums / n: exact
    assert np.array_equal(res["colour"], ref["colour"])
    assert np.array_equal(res["pixels"], ref["pixels"])          # the integer waterfall pixel row
    assert np.array_equal(_sc(res), ref["scalars"])
    bank.close()


@pytest.mark.parametrize("N,window", [(256, True), (512, False), (1024, True), (1024, False), (2048, True),
                                      (4096, True), (8192, False), (16384, True), (16384, False),
                                      (32768, True), (65536, True), (65536, False)])
def test_waterfall_rounding_noise_is_bit_exact(ssdr, N, window):
    """Adversarial for last-bit differences: one strong tone on an EXACT bin (optionally a second weak
    one) and no noise.  Every other bin then holds pure float32 rounding noise, which is reproduced only
    if every operation (and every fused / unfused rounding) happens exactly as the spec states."""
    B, n = 12, 2
    rng = np.random.default_rng(N)
    t = np.arange(n * N)
    iq = np.empty((B, n, N), np.complex64)
    for b in range(B):
        k1, k2 = rng.integers(0, N, 2)
        x = 0.5 * np.exp(2j * np.pi * (k1 * t / N + rng.uniform()))
        if b % 3 == 1:
            x = x + 1e-3 * np.exp(2j * np.pi * (k2 * t / N + rng.uniform()))
        if b % 3 == 2:
            x = np.rint(x * 32768) / 32768                       # integer counts, as a real int16 stream
        iq[b] = (x * 32768).astype(np.complex64).reshape(n, N)
    bank = ssdr.WaterfallBank(N, B, n, window=window)
    res = bank.process(iq)
    ref = c_oracle.wf_rows(iq, window=window, threads=8)
    assert np.array_equal(res["spectrum"], ref["spectrum"])
    assert np.array_equal(res["pixels"], ref["pixels"])
    bank.close()


def test_config1_single_1024_frame_golden_and_float64(ssdr):
    """BASELINE config 1: one 1024-pt frame.  Golden fixture + boundary-aware check vs float64."""
    g = np.load(os.path.join(GOLD, "tier_u_waterfall.npz"))
    bank = ssdr.WaterfallBank(1024, 1, 1)
    res = bank.process(g["iq1"][None])
    assert np.array_equal(res["pixels"], g["pixels1"]) and np.array_equal(res["spectrum"], g["spectrum1"])
    by = res["spectrum"][0].astype(np.uint8)                     # n_avg = 1: spectrum is the byte line
    assert np.array_equal(by, g["bytes1"])
    mism, unexplained = tier_u.compare_bytes_boundary_aware(by, g["iq1"][0])
    assert unexplained == 0 and mism <= 2
    bank.close()
    b2 = ssdr.WaterfallBank(16384, 2, 2)
    b2.set_display(zoom=3)
    r2 = b2.process(g["iq2"])
    assert np.array_equal(r2["pixels"], g["pixels2"]) and np.array_equal(_sc(r2), g["scalars2"])
    b2.close()


def test_wire_format_s16be_input(ssdr):
    """K6 fused into the FFT load: big-endian int16 I,Q (kiwi/client.py:449-453)."""
    for N, B, n in ((1024, 4, 2), (16384, 2, 1)):
        iq = tier_u.synth_batch(B, n, N, seed=3, quantise=True)
        q = np.stack([iq.real, iq.imag], -1).astype(">i2")
        wire = np.ascontiguousarray(q).view(np.uint8).reshape(B, n, N, 4)
        bank = ssdr.WaterfallBank(N, B, n)
        r_wire = bank.process(wire)
        r_cf = bank.process(iq)
        ref = c_oracle.wf_rows(iq)
        assert np.array_equal(r_wire["pixels"], ref["pixels"]) and np.array_equal(r_cf["pixels"], ref["pixels"])
        bank.close()
    raw = np.random.default_rng(0).integers(-32768, 32768, 4096).astype(">i2")
    want = (raw[0::2].astype(np.float32) + 1j * raw[1::2].astype(np.float32)).astype(np.complex64)
    assert np.array_equal(ssdr.unpack_iq(raw.tobytes()), want)


def test_colorrow_tier_p_reference_golden(ssdr):
    """The reference's own input (finished uint8 W/F lines) against the fixtures the UNMODIFIED
    reference produced: averaging utils:881-886 + spectrum_db2col utils:787-813, bit-exact."""
    g = np.load(os.path.join(GOLD, "tier_p_waterfall.npz"))
    for i in range(int(g["n_cases"])):
        k = "c%02d_" % i
        lines = g[k + "lines"]
        n, W = lines.shape
        bank = ssdr.WaterfallBank(W, 1, n)
        bank.set_display(zoom=int(g[k + "zoom"]), auto_scale=bool(g[k + "auto"]), delta_low_db=int(g[k + "dlow"]),
                         delta_high_db=int(g[k + "dhigh"]), low_clip_db=-120.0, dynamic_range=40.0)
        res = bank.colorrow(lines[None])
        assert np.array_equal(res["spectrum"][0], g[k + "spectrum"]), i
        assert np.array_equal(res["colour"][0], g[k + "colour"]), i
        assert np.array_equal(res["pixels"][0], np.rint(g[k + "colour"]).astype(np.uint8)), i
        assert np.array_equal(_sc(res)[0][[0, 2, 3, 4]], g[k + "scalars"][[0, 2, 3, 4]]), i
        bank.close()


def test_colorrow_batched_random(ssdr):
    rng = np.random.default_rng(11)
    for W, B, n in ((1024, 37, 10), (256, 40, 3), (16384, 3, 100), (2048, 9, 1)):
        lines = np.clip(rng.normal(120, 8, (B, n, W)), 0, 255).astype(np.uint8)
        lines[0] = 0                                  # empty band: all-zero lines
        lines[1 % B] = 255                            # saturated
        bank = ssdr.WaterfallBank(W, B, n)
        bank.set_display(zoom=5, delta_low_db=-3, delta_high_db=4)
        res = bank.colorrow(lines)
        for b in range(B):
            st = tier_p.ColourState()
            st.zoom, st.delta_low_db, st.delta_high_db = 5, -3, 4
            spec, col, px = tier_p.waterfall_line(lines[b], st)
            assert np.array_equal(spec, res["spectrum"][b]) and np.array_equal(col, res["colour"][b])
            assert np.array_equal(px, res["pixels"][b])
            assert np.float32(st.low_clip_db) == res["scalars"]["low_clip_db"][b]
        bank.close()


def test_auto_scale_off_keeps_last_levels(ssdr):
    """supersdr.py:408-410: with wf_auto_scaling off, low_clip_db / dynamic_range keep their values."""
    rng = np.random.default_rng(2)
    lines = np.clip(rng.normal(90, 5, (1, 4, 1024)), 0, 255).astype(np.uint8)
    bank = ssdr.WaterfallBank(1024, 1, 4)
    r_auto = bank.colorrow(lines)
    low, dyn = float(r_auto["scalars"]["low_clip_db"][0]), float(r_auto["scalars"]["dynamic_range"][0])
    bank.set_display(auto_scale=False, low_clip_db=low, dynamic_range=dyn, delta_low_db=2)
    r_man = bank.colorrow(np.clip(lines + 20, 0, 255).astype(np.uint8))
    st = tier_p.ColourState()
    st.wf_auto_scaling, st.low_clip_db, st.dynamic_range, st.delta_low_db = False, np.float32(low), np.float32(dyn), 2
    _, col, _ = tier_p.waterfall_line(np.clip(lines[0] + 20, 0, 255).astype(np.uint8), st)
    assert np.array_equal(col, r_man["colour"][0])
    bank.close()


def test_edge_inputs(ssdr):
    """Silence, full-scale tone at an exact bin, DC, and a saturated square-ish input."""
    N = 1024
    n = np.arange(N)
    frames = np.stack([np.zeros(N), 32768 * np.exp(2j * np.pi * 100 * n / N), np.full(N, 1000 + 0j),
                       32767 * np.sign(np.sin(2 * np.pi * 5 * n / N)) * (1 + 1j)]).astype(np.complex64)
    bank = ssdr.WaterfallBank(N, 4, 1)
    res = bank.process(frames[:, None, :])
    ref = c_oracle.wf_rows(frames[:, None, :])
    assert np.array_equal(res["pixels"], ref["pixels"]) and np.array_equal(res["spectrum"], ref["spectrum"])
    assert res["spectrum"][0].max() == 0                         # silence -> byte 0
    assert res["spectrum"][1][512 + 100] == 245                  # 0 dBFS tone -> -10 dBm -> byte 245
    assert res["spectrum"][2].argmax() == 512                    # DC lands in the centre bin (fftshift)
    bank.close()


@pytest.mark.parametrize("N,B,n", [(16384, 2, 100), (8192, 2, 100), (8192, 2, 33), (65536, 1, 100), (2048, 3, 100)])
def test_averaging_extremes_bit_exact(ssdr, N, B, n):
    """averaging_n up to its cap of 100 (supersdr.py:376-385): 15-bit sums, the histogram rank selection where it
    fits the frame buffer and the bisection where it does not, against the C oracle."""
    iq = tier_u.synth_batch(B, n, N, seed=7 * N + n)
    bank = ssdr.WaterfallBank(N, B, n)
    res = bank.process(iq)
    ref = c_oracle.wf_rows(iq, threads=8)
    assert np.array_equal(res["spectrum"], ref["spectrum"])
    assert np.array_equal(res["colour"], ref["colour"])
    assert np.array_equal(res["pixels"], ref["pixels"])
    assert np.array_equal(_sc(res), ref["scalars"])
    bank.close()


def test_constant_and_two_level_rows(ssdr):
    """Rows whose keys are all equal (digital silence) or take two values only: the degenerate inputs of the
    histogram rank selection, through the FFT path and through the uint8-line entry."""
    N = 16384
    iq = np.zeros((2, 2, N), np.complex64)
    iq[1, :, :] = 3000.0                                   # DC only: one strong bin (+ window leakage), silence elsewhere
    bank = ssdr.WaterfallBank(N, 2, 2)
    res = bank.process(iq)
    ref = c_oracle.wf_rows(iq)
    for k in ("spectrum", "colour", "pixels"):
        assert np.array_equal(res[k], ref[k])
    assert np.array_equal(_sc(res), ref["scalars"])
    lines = np.full((2, 3, N), 77, np.uint8)
    lines[1, :, ::2] = 200
    got = bank3 = None
    bank3 = ssdr.WaterfallBank(N, 2, 3)
    got = bank3.colorrow(lines)
    for b in range(2):
        st = tier_p.ColourState()
        spec, col, px = tier_p.waterfall_line(lines[b], st)
        assert np.array_equal(got["spectrum"][b], spec) and np.array_equal(got["colour"][b], col)
        assert np.array_equal(got["pixels"][b], px)
    bank.close(); bank3.close()


def test_bad_arguments_raise(ssdr):
    with pytest.raises(ssdr.SsdrError):
        ssdr.WaterfallBank(1000, 1, 1)            # not a power of two
    with pytest.raises(ssdr.SsdrError):
        ssdr.WaterfallBank(1024, 1, 101)          # averaging_n capped at 100 (supersdr.py:378)
    with pytest.raises(ssdr.SsdrError):
        ssdr.WaterfallBank(131072, 1, 1)          # largest supported frame is 65536
    big = ssdr.WaterfallBank(32768, 1, 1)
    with pytest.raises(ssdr.SsdrError):
        big.colorrow(np.zeros((1, 1, 32768), np.uint8))      # the uint8-line entry stops at 16384
    big.close()
    bank = ssdr.WaterfallBank(1024, 2, 1)
    with pytest.raises(ValueError):
        bank.process(np.zeros((1, 1, 1024), np.complex64))
    with pytest.raises(ssdr.SsdrError):
        bank.set_display(first=1, count=5)
    bank.close()


def test_full_size_config2_properties(ssdr):
    """BASELINE config 2 at full size (4096 ch x 10 x 16384, generated in HBM): sampled channels are
    bit-exact against the oracle on the very same bits; size-independent properties hold for all rows."""
    B, n, N = 4096, 10, 16384
    iq = ssdr.DeviceBuffer(B * n * N * 8)
    px = ssdr.DeviceBuffer(B * N)
    ssdr._lib.check(ssdr.lib.ssdr_synth_iq_dev(iq.ptr, ssdr.SSDR_IQ_CF32, B, n, N, 1234))
    bank = ssdr.WaterfallBank(N, B, n)
    bank.process_dev(iq.ptr, ssdr.SSDR_IQ_CF32, px.ptr)
    bank.sync()
    pix = px.download(np.uint8, (B, N))
    for ch in (0, 147, 148, 2049, 4095):
        x = iq.download(np.complex64, (1, n, N), offset_bytes=ch * n * N * 8)
        assert np.array_equal(c_oracle.wf_rows(x)["pixels"][0], pix[ch]), ch
    assert pix.max() <= 254                                       # colour row range [0, 254]
    assert np.all(pix.max(axis=1) == 254)                         # p100 maps to 254 in every row (dyn >= 40 from tones)
    frac0 = (pix == 0).mean(axis=1)
    assert np.all((frac0 > 0.2) & (frac0 < 0.6))                  # ~40 % of bins sit at/below the 40th percentile
    # idempotence: same input, same rows; and channel independence: a sub-batch gives the same rows
    bank.process_dev(iq.ptr, ssdr.SSDR_IQ_CF32, px.ptr)
    bank.sync()
    assert np.array_equal(px.download(np.uint8, (B, N)), pix)
    sub = ssdr.WaterfallBank(N, 8, n)
    px2 = ssdr.DeviceBuffer(8 * N)
    off = 1000 * n * N * 8
    import ctypes
    sub.process_dev(ctypes.c_void_p(iq.ptr.value + off), ssdr.SSDR_IQ_CF32, px2.ptr)
    sub.sync()
    assert np.array_equal(px2.download(np.uint8, (8, N)), pix[1000:1008])
    for o in (bank, sub):
        o.close()
    for o in (iq, px, px2):
        o.free()


def test_large_frames_fused_kernel_equals_three_kernel_path(ssdr, tmp_path):
    """The opt-in fused large-frame kernel (SSDR_WF_BIG=fused: front pass + sub-transforms on one SM, sub-frames through an
    L2-resident scratch, byte sums in tensor memory) gives the rows of the default three-kernel path bit for bit.  The
    switch is read once per process, so the fused run is a subprocess."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "fused.npz"
    code = ("import sys, numpy as np; sys.path.insert(0, %r)\n"
            "import supersdr_b200 as S\nfrom oracle import tier_u\nres = {}\n"
            "for N, B, n in ((32768, 151, 2), (65536, 5, 3)):\n"
            "    bank = S.WaterfallBank(N, B, n); bank.set_display(zoom=1)\n"
            "    r = bank.process(tier_u.synth_batch(B, n, N, seed=N + B))\n"
            "    res['px%%d' %% N], res['sp%%d' %% N] = r['pixels'], r['spectrum']\n"
            "np.savez(%r, **res)\n") % (root, str(out))
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, SSDR_WF_BIG="fused"), timeout=600)
    got = np.load(out)
    for N, B, n in ((32768, 151, 2), (65536, 5, 3)):
        bank = ssdr.WaterfallBank(N, B, n)
        bank.set_display(zoom=1)
        r = bank.process(tier_u.synth_batch(B, n, N, seed=N + B))
        assert np.array_equal(got["px%d" % N], r["pixels"]) and np.array_equal(got["sp%d" % N], r["spectrum"])
        bank.close()


@pytest.mark.gpu
@pytest.mark.parametrize("N", [16384, 8192, 4096, 2048, 1024, 512])
@pytest.mark.parametrize("fmt", ["cf32", "s16be"])
def test_staged_kernel_equals_direct_load_kernel(ssdr, fmt, N):
    """512- to 16384-point frames take the TMA-staged kernel (one tensor-map tile per warp and frame) when the
    input is local, the direct-load kernel when it is flagged as peer (NVLink) input.  Same arithmetic routines, so every
    output must be bit-identical -- on more channels than CTAs, so that a CTA walks several channels (the tile of the next
    channel is issued behind the row stage).  An input pointer that is not 16-byte aligned is refused (header contract)."""
    import ctypes
    B, n = min(300 * (16384 // N), 2500) + 3, 3          # + 3: the last CTA's frame groups are not all busy
    code = ssdr.SSDR_IQ_CF32 if fmt == "cf32" else ssdr.SSDR_IQ_S16BE
    sb = 8 if fmt == "cf32" else 4
    iq = ssdr.DeviceBuffer(B * n * N * sb + 64)
    ssdr._lib.check(ssdr.lib.ssdr_synth_iq_dev(iq.ptr, code, B, n, N, 4321))
    outs = []
    for remote in (False, True):
        bank = ssdr.WaterfallBank(N, B, n)
        bank.set_remote_input(remote)
        px, col, spec = ssdr.DeviceBuffer(B * N), ssdr.DeviceBuffer(B * N * 4), ssdr.DeviceBuffer(B * N * 4)
        bank.process_dev(iq.ptr, code, px.ptr, col.ptr, spec.ptr)
        bank.sync()
        outs.append((px.download(np.uint8, (B, N)), col.download(np.float32, (B, N)), spec.download(np.float32, (B, N))))
        if not remote:
            with pytest.raises(ssdr.SsdrError):
                bank.process_dev(ctypes.c_void_p(iq.ptr.value + 8), code, px.ptr)
        bank.close()
        for o in (px, col, spec):
            o.free()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    if fmt == "cf32":
        x = iq.download(np.complex64, (2, n, N), offset_bytes=150 * n * N * sb)
    else:
        raw = iq.download(np.uint8, (2 * n * N * 4,), offset_bytes=150 * n * N * sb).view(">i2").astype(np.float32)
        x = (raw[0::2] + 1j * raw[1::2]).astype(np.complex64).reshape(2, n, N)
    assert np.array_equal(c_oracle.wf_rows(x)["pixels"], outs[0][0][150:152])
    iq.free()
