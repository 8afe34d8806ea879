"""GPU parity tests (through the C ABI) for the demodulator (Tier U, 1e-5 RMS vs the float64
oracle) and the audio interpolator (Tier P, against fixtures from the unmodified reference)."""
import os

import numpy as np
import pytest

from oracle import tier_p, tier_u

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
RMS_TOL = 1e-5          # north_star: float32 demod output within 1e-5 RMS of the numpy path


@pytest.fixture(params=["ffma", "tcgen05"])
def engine(request):
    """Every demodulator test runs on both FIR engines of the fused kernel (ssdr_demod_set_engine): the handle reads its
    default engine from SSDR_DEMOD_ENGINE when it is created."""
    old = os.environ.get("SSDR_DEMOD_ENGINE")
    os.environ["SSDR_DEMOD_ENGINE"] = request.param
    yield request.param
    if old is None:
        del os.environ["SSDR_DEMOD_ENGINE"]
    else:
        os.environ["SSDR_DEMOD_ENGINE"] = old


def _rel_rms(a, b):
    return np.sqrt(np.mean((a - b) ** 2)) / max(np.sqrt(np.mean(b ** 2)), 1e-30)


@pytest.mark.parametrize("mode", ["usb", "lsb", "cw", "am", "nbfm"])
@pytest.mark.parametrize("hang,on,slope", [(False, True, 0), (True, True, 6), (False, False, 0)])
def test_demod_vs_float64_oracle(ssdr, engine, mode, hang, on, slope):
    B, n = 6, 512 * 12
    kw = dict(mode=mode, hang=hang, on=on, slope=slope, decay=1000 if mode == "cw" else 4000)
    bank = ssdr.DemodBank(B, n)
    bank.set_all(**kw)
    iq = np.stack([tier_u.synth_demod_iq(mode, n, seed=10 + b, level=0.1 / (b + 1)) for b in range(B)])
    r1 = bank.process(iq[:, : n // 2].copy())          # streamed in two calls: state carries over
    r2 = bank.process(iq[:, n // 2:].copy())
    got = np.concatenate([r1["pcm_f32"], r2["pcm_f32"]], 1)
    gi = np.concatenate([r1["pcm_i16"], r2["pcm_i16"]], 1)
    rssi = np.concatenate([r1["rssi"], r2["rssi"]], 1)
    for b in range(B):
        ref, rr = tier_u.demod(iq[b], tier_u.DemodParams(**kw), tier_u.DemodState())
        assert _rel_rms(got[b], ref) < RMS_TOL
        assert np.abs(gi[b].astype(int) - tier_u.pcm_to_i16(ref).astype(int)).max() <= 1
        assert np.abs(rssi[b] - rr).max() < 1e-3
    bank.close()


def test_demod_golden_fixture(ssdr, engine):
    g = np.load(os.path.join(GOLD, "tier_u_demod.npz"))
    for mode in tier_u.MODES:
        bank = ssdr.DemodBank(1, 4096)
        bank.set_all(mode=mode, decay=1000 if mode == "cw" else 4000, hang=(mode == "cw"))
        got = bank.process(g["iq_" + mode][None])
        assert _rel_rms(got["pcm_f32"][0], g["pcm_" + mode].astype(np.float64)) < RMS_TOL
        assert np.abs(got["rssi"][0] - g["rssi_" + mode]).max() < 1e-3
        bank.close()


def test_demod_mixed_modes_and_chunking(ssdr, engine):
    """Config 4 in miniature: modes by ch % 5, per-channel tuning offsets; and the result does not
    depend on how the stream is cut into calls (per-frame state is bit-reproducible)."""
    modes = ["am", "lsb", "usb", "cw", "nbfm"]
    B, n = 25, 512 * 8
    params = [ssdr.demod_params(modes[b % 5], f_off=100.0 * (b % 3), thresh=-80 - b, decay=400 + 300 * b) for b in range(B)]
    iq = np.stack([tier_u.synth_demod_iq(modes[b % 5], n, seed=b) for b in range(B)])
    one = ssdr.DemodBank(B, n)
    one.set_params(0, params)
    whole = one.process(iq)
    cut = ssdr.DemodBank(B, n)
    cut.set_params(0, params)
    parts = [cut.process(iq[:, i:i + 512 * k].copy()) for i, k in ((0, 1), (512, 3), (2048, 4))]
    assert np.array_equal(np.concatenate([p["pcm_f32"] for p in parts], 1), whole["pcm_f32"])
    assert np.array_equal(np.concatenate([p["pcm_i16"] for p in parts], 1), whole["pcm_i16"])
    for b in range(B):
        p = tier_u.DemodParams(modes[b % 5], f_off=100.0 * (b % 3), thresh=-80 - b, decay=400 + 300 * b)
        ref, _ = tier_u.demod(iq[b], p, tier_u.DemodState())
        assert _rel_rms(whole["pcm_f32"][b], ref) < RMS_TOL, b
    one.close(); cut.close()


def test_demod_host_pipeline_time_blocks(ssdr, engine):
    """A call large enough to be cut into several time blocks by the host pipeline (ssdr_demod_process: > 64 MiB of
    input) gives, per channel, exactly what one small call gives, into caller-owned output buffers."""
    modes = ["usb", "am", "nbfm", "cw"]
    B, n = 640, 512 * 48                                   # 126 MB of complex64 -> 2 blocks
    src = np.stack([tier_u.synth_demod_iq(m, n, seed=21 + i) for i, m in enumerate(modes)])
    iq = np.ascontiguousarray(src[np.arange(B) % 4])
    params = [ssdr.demod_params(modes[b % 4], hang=(b % 4 == 3)) for b in range(B)]
    big = ssdr.DemodBank(B, n)
    big.set_params(0, params)
    out = {"pcm_f32": np.empty((B, n), np.float32), "pcm_i16": np.empty((B, n), np.int16), "rssi": np.empty((B, n // 512), np.float32)}
    res = big.process(iq, out=out)
    assert res["pcm_f32"] is out["pcm_f32"] and res["rssi"] is out["rssi"]
    small = ssdr.DemodBank(4, n)
    small.set_params(0, params[:4])
    ref = small.process(src)
    for k in ("pcm_f32", "pcm_i16", "rssi"):
        assert np.array_equal(res[k][:4], ref[k])
        assert np.array_equal(res[k], res[k][np.arange(B) % 4])          # replicas of a channel are identical
    o, _ = tier_u.demod(src[0], tier_u.DemodParams("usb"), tier_u.DemodState())
    assert _rel_rms(res["pcm_f32"][B - 4], o) < RMS_TOL
    with pytest.raises(ValueError):
        big.process(iq, out={"pcm_f32": np.empty((B, n), np.float64)})
    big.close(); small.close()


def test_demod_wire_format_and_sideband_rejection(ssdr, engine):
    n = 512 * 16
    iq = tier_u.synth_demod_iq("usb", n, seed=4)
    iq = (np.rint(iq.real) + 1j * np.rint(iq.imag)).astype(np.complex64)
    wire = np.ascontiguousarray(np.stack([iq.real, iq.imag], -1).astype(">i2")).view(np.uint8).reshape(1, n, 4)
    bank = ssdr.DemodBank(1, n)
    a = bank.process(iq[None])["pcm_f32"][0]
    bank.reset()
    b = bank.process(wire)["pcm_f32"][0]
    assert np.array_equal(a, b)
    F = np.abs(np.fft.rfft(a[-4096:] * np.hanning(4096)))
    f = np.fft.rfftfreq(4096, 1 / 12000)
    assert abs(f[F.argmax()] - 1000) < 5             # +1 kHz tone demodulated, -1 kHz image rejected:
    bank.set_all(mode="lsb"); bank.reset()
    l = bank.process(iq[None])["pcm_f32"][0]
    assert np.sqrt(np.mean(l[-4096:] ** 2)) > 0.5 * np.sqrt(np.mean(a[-4096:] ** 2))   # LSB hears the -1 kHz tone
    bank.close()


def test_demod_silence_and_full_scale(ssdr, engine):
    bank = ssdr.DemodBank(2, 1024)
    x = np.zeros((2, 1024), np.complex64)
    x[1] = 32767 * np.exp(2j * np.pi * 1000 * np.arange(1024) / 12000)
    r = bank.process(x)
    assert np.all(r["pcm_f32"][0] == 0) and np.all(r["pcm_i16"][0] == 0) and np.all(np.isfinite(r["rssi"]))
    ref, rr = tier_u.demod(x[1], tier_u.DemodParams("usb"), tier_u.DemodState())
    assert _rel_rms(r["pcm_f32"][1], ref) < RMS_TOL and abs(r["rssi"][1, -1] - rr[-1]) < 1e-3
    with pytest.raises(ssdr.SsdrError):
        bank.process(np.zeros((2, 500), np.complex64))       # not a multiple of 512
    bank.close()


def test_demod_engines_agree_and_interleave(ssdr):
    """The two engines share the per-channel state format: a stream processed alternately by one and the other stays
    within the tolerance of the float64 oracle; channels with distinct filters (quads of one) and a batch that is not a
    multiple of four are covered."""
    modes = ["usb", "am", "cw", "nbfm", "lsb", "usb", "am"]
    B, n = len(modes), 512 * 8
    params = [ssdr.demod_params(m, hc=2700.0 + 100 * b) if m == "usb" else ssdr.demod_params(m) for b, m in enumerate(modes)]
    iq = np.stack([tier_u.synth_demod_iq(m, n, seed=40 + b) for b, m in enumerate(modes)])
    bank = ssdr.DemodBank(B, n)
    bank.set_params(0, params)
    parts = []
    for i, eng in enumerate(["tcgen05", "ffma", "tcgen05", "tcgen05"]):
        bank.set_engine(eng)
        parts.append(bank.process(iq[:, i * 1024:(i + 1) * 1024].copy())["pcm_f32"])
    got = np.concatenate(parts, 1)
    for b, m in enumerate(modes):
        p = tier_u.DemodParams(m, hc=2700.0 + 100 * b) if m == "usb" else tier_u.DemodParams(m)
        ref, _ = tier_u.demod(iq[b], p, tier_u.DemodState())
        assert _rel_rms(got[b], ref) < RMS_TOL, (b, m)
    bank.close()


def test_demod_auto_engine_picks_by_tile_fill(ssdr):
    """The default engine ("auto") is the tcgen05 kernel when the tensor-core tiles are at least half full (channels that
    share a filter) and the FFMA kernel otherwise -- bit for bit the explicit engine's output in both cases."""
    n = 512 * 4
    for params, expect in (([ssdr.demod_params("usb")] * 8, "tcgen05"),
                           ([ssdr.demod_params("usb", hc=2500.0 + 100 * b) for b in range(3)], "ffma")):
        B = len(params)
        iq = np.stack([tier_u.synth_demod_iq("usb", n, seed=70 + b) for b in range(B)])
        out = {}
        for eng in ("auto", "ffma", "tcgen05"):
            bank = ssdr.DemodBank(B, n, engine=eng)
            bank.set_params(0, params)
            out[eng] = bank.process(iq)["pcm_f32"]
            bank.close()
        other = "ffma" if expect == "tcgen05" else "tcgen05"
        assert np.array_equal(out["auto"], out[expect])
        assert not np.array_equal(out["auto"], out[other])          # the engines differ in the last bits
        assert _rel_rms(out["ffma"].astype(np.float64), out["tcgen05"].astype(np.float64)) < RMS_TOL


def test_demod_auto_engine_many_filters_cost_model(ssdr):
    """More rounds than SMs (the tcgen05 engine shares one filter per CTA round): with two channels per filter the
    measured cost model of AUTO (capi.cu: demod_auto_prefers_tc, scripts/demod_hetero.py) picks the FFMA engine although
    the tiles are half full; with one shared filter it stays on the tensor cores.  Bit for bit the explicit engine."""
    n = 512 * 8
    for per, expect in ((2, "ffma"), (600, "tcgen05")):
        B = 600
        uniq = [ssdr.demod_params("usb", hc=2400.0 + g) for g in range(B // per)]
        params = [uniq[b // per] for b in range(B)]
        iq = np.stack([tier_u.synth_demod_iq("usb", n, seed=300 + b % 7) for b in range(B)])
        out = {}
        for eng in ("auto", expect):
            bank = ssdr.DemodBank(B, n, engine=eng)
            bank.set_params(0, params)
            out[eng] = bank.process(iq)["pcm_f32"]
            bank.close()
        assert np.array_equal(out["auto"], out[expect]), per
        ref, _ = tier_u.demod(iq[B - 1], tier_u.DemodParams("usb", hc=2400.0 + (B - 1) // per), tier_u.DemodState())
        assert _rel_rms(out["auto"][B - 1], ref) < RMS_TOL


def test_interp_reference_golden(ssdr):
    """kiwi_sound.play_buffer (utils_supersdr.py:1121-1138) fixtures from the unmodified reference:
    int16 stereo within 1 LSB (np.convolve's summation order is BLAS-dependent)."""
    g = np.load(os.path.join(GOLD, "tier_p_audio.npz"))
    ib = ssdr.InterpBank(1, 4, taps=g["h"], max_samples=512)
    st = tier_p.InterpState()
    exact = 0
    for k in range(g["x"].shape[0]):
        out, mono = ib.process(g["x"][k][None], float(g["volume"][k]), float(g["balance"][k]), want_mono=True)
        d = np.abs(out[0].astype(int) - g["out"][k].astype(int))
        # the int16 wrap of numpy's astype is reproduced (block 3 is full-scale DC at volume 150)
        assert (np.minimum(d, 65536 - d)).max() <= 1
        exact += int(d.max() == 0)
        buf, _ = tier_p.play_buffer(g["x"][k], st, int(g["volume"][k]), float(g["balance"][k]))
        assert _rel_rms(mono[0], buf) < RMS_TOL or np.abs(buf).max() == 0
    assert exact >= 8
    ib.close()


def test_interp_batched_streaming(ssdr):
    rng = np.random.default_rng(1)
    B, n = 9, 512
    ib = ssdr.InterpBank(B, 4, max_samples=2 * n)
    sts = [tier_p.InterpState() for _ in range(B)]
    for it in range(5):
        nn = n if it % 2 == 0 else 2 * n
        x = rng.integers(-32768, 32768, (B, nn)).astype(np.int16)
        vol = rng.integers(0, 16, B) * 10.0
        bal = np.round(rng.uniform(-1, 1, B), 2).astype(np.float32)
        out = ib.process(x, vol, bal)
        for b in range(B):
            _, o2 = tier_p.play_buffer(x[b], sts[b], float(vol[b]), float(bal[b]))
            d = np.abs(out[b].astype(int) - o2.astype(int))
            assert np.minimum(d, 65536 - d).max() <= 1
    ib.close()
    f = ssdr.filtering(6000, 48000)
    x = rng.standard_normal(5000)
    assert np.abs(f.lowpass(x) - np.convolve(x, f.h, "valid")).max() < 1e-12


def test_resample_line_reference_golden(ssdr):
    """play_buffer's non-integer-ratio path (utils_supersdr.py:1125-1126) against fixtures from the unmodified
    reference (20.25 kHz -> 48 kHz, resample_poly(x, 64, 27, padtype="line")[:-1])."""
    g = np.load(os.path.join(GOLD, "tier_p_audio_resample.npz"))
    rs = ssdr.ResampleLine(int(g["up"]), int(g["down"]))
    for k in range(g["x"].shape[0]):
        out = rs.process(g["x"][k][None], float(g["volume"][k]), float(g["balance"][k]))[0]
        assert out.shape == g["out"][k].shape
        d = np.abs(out.astype(int) - g["out"][k].astype(int))
        assert (np.minimum(d, 65536 - d)).max() <= 1           # float64 summation order of upfirdn is not pinned: <= 1 LSB


@pytest.mark.parametrize("up,down,n", [(64, 27, 512), (64, 27, 1024), (160, 147, 441), (3, 2, 100), (5, 7, 333)])
def test_resample_line_vs_scipy(ssdr, up, down, n):
    from scipy.signal import resample_poly
    rng = np.random.default_rng(up * 1000 + n)
    B = 4
    x = rng.integers(-30000, 30000, (B, n)).astype(np.int16)
    x[1] = (-25000 + (50000 // n) * np.arange(n)).astype(np.int16)         # a ramp: the "line" extension is exercised
    vol = np.array([100, 50, 70, 130], np.float32)
    rs = ssdr.ResampleLine(up, down)
    out, mono = rs.process(x, vol, 0.25, want_mono=True)
    for b in range(B):
        ref = resample_poly(x[b].astype(np.float64) * (float(vol[b]) / 100), up, down, padtype="line")[:-1]
        assert mono[b].shape == ref.shape
        assert np.abs(mono[b] - ref).max() <= 1e-9 * max(1.0, np.abs(ref).max())


def test_adpcm_reference_golden(ssdr):
    """IMA-ADPCM decode against fixtures from the reference's ImaAdpcmDecoder (kiwi/client.py:58-87): four streams in
    one launch, streamed in three calls (state carried), including the clamped-index and saturating cases."""
    g = np.load(os.path.join(GOLD, "adpcm.npz"))
    data, pcm, states = g["data"], g["pcm"], g["states"]
    dec = ssdr.ImaAdpcmDecoder(batch=data.shape[0])
    for c in range(3):
        out = dec.decode(data[:, c * 333:(c + 1) * 333])
        assert np.array_equal(out, pcm[:, c * 666:(c + 1) * 666])
        assert np.array_equal(dec.state, states[:, c])
    one = ssdr.ImaAdpcmDecoder()                       # the reference's single-stream surface: bytes in, samples out
    assert np.array_equal(one.decode(bytes(data[0, :333])), pcm[0, :666]) and (one.index, one.prev) == tuple(states[0, 0])
    big = ssdr.ImaAdpcmDecoder(batch=300)              # aligned 16-byte path + tail, many streams
    d = np.tile(data[0, :512 + 5], (300, 1))
    ref = ssdr.ImaAdpcmDecoder().decode(bytes(data[0, :512 + 5]))
    assert np.array_equal(big.decode(d), np.tile(ref, (300, 1)))


def test_argument_validation_of_the_advice_items(ssdr):
    """an AGC decay that would overflow the float32 envelope scan is refused (instead of poisoning the
    channel's state with NaN); the *_dev entry points refuse device pointers that are not 16-byte aligned (instead of
    faulting); batches beyond the grid's y dimension are refused with a message."""
    import ctypes
    bank = ssdr.DemodBank(2, 1024)
    with pytest.raises(ssdr.SsdrError):
        bank.set_params(0, [ssdr.demod_params("usb", decay=0.5)] * 2)
    bank.set_params(0, [ssdr.demod_params("usb", decay=1.0)] * 2)            # the shortest accepted decay stays finite
    x = np.stack([tier_u.synth_demod_iq("usb", 1024, seed=s) for s in (1, 2)])
    r = bank.process(x)
    assert np.all(np.isfinite(r["pcm_f32"])) and np.all(np.isfinite(bank.process(x)["pcm_f32"]))
    iq = ssdr.DeviceBuffer(2 * 1024 * 8 + 64)
    out = ssdr.DeviceBuffer(2 * 1024 * 4 + 64)
    with pytest.raises(ssdr.SsdrError):
        bank.process_dev(ctypes.c_void_p(iq.ptr.value + 8), ssdr.SSDR_IQ_CF32, 1024, out.ptr)
    with pytest.raises(ssdr.SsdrError):
        bank.process_dev(iq.ptr, ssdr.SSDR_IQ_CF32, 1024, ctypes.c_void_p(out.ptr.value + 4))
    bank.process_dev(iq.ptr, ssdr.SSDR_IQ_CF32, 1024, out.ptr)
    bank.sync()
    bank.close(); iq.free(); out.free()
    with pytest.raises(ssdr.SsdrError):
        ssdr.InterpBank(70000, 4)
