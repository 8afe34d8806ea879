"""CPU: the Tier-P restatement (oracle/tier_p.py, oracle/c/ssdr_oracle.c) against the committed
golden vectors that the UNMODIFIED reference produced (oracle/make_golden.py), and -- when
/root/reference is present (build container) -- against the imported reference itself."""
import os
import queue

import numpy as np
import pytest

from oracle import tier_p, c_oracle, ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cases():
    g = np.load(os.path.join(GOLD, "tier_p_waterfall.npz"))
    for i in range(int(g["n_cases"])):
        k = "c%02d_" % i
        yield {n: g[k + n] for n in ("lines", "zoom", "auto", "dlow", "dhigh", "spectrum", "colour", "scalars")}


def _state(c):
    st = tier_p.ColourState()
    st.zoom, st.wf_auto_scaling = int(c["zoom"]), bool(c["auto"])
    st.delta_low_db, st.delta_high_db = int(c["dlow"]), int(c["dhigh"])
    return st


def test_waterfall_restatement_matches_reference_golden():
    n = 0
    for c in _cases():
        st = _state(c)
        spec, col, px = tier_p.waterfall_line(c["lines"], st)
        assert spec.dtype == np.float32 and np.array_equal(spec, c["spectrum"])
        assert col.dtype == np.float32 and np.array_equal(col, c["colour"])
        got = np.array([st.low_clip_db, st.high_clip_db, st.dynamic_range, st.wf_min_db, st.wf_max_db], np.float32)
        if st.wf_auto_scaling:
            assert np.array_equal(got, c["scalars"])
        else:
            assert np.array_equal(got[[0, 2, 3, 4]], c["scalars"][[0, 2, 3, 4]])
        assert np.array_equal(px, np.rint(c["colour"]).astype(np.uint8))
        n += 1
    assert n == 21


def test_c_colour_row_matches_reference_golden():
    for c in _cases():
        lines = c["lines"]
        sums = lines.astype(np.uint16).sum(axis=0).astype(np.uint16)
        spec, col, px, sc = c_oracle.colour_row(sums, lines.shape[0], zoom=int(c["zoom"]), auto_scale=bool(c["auto"]),
                                                delta_low_db=int(c["dlow"]), delta_high_db=int(c["dhigh"]))
        assert np.array_equal(spec, c["spectrum"])
        assert np.array_equal(col, c["colour"])
        assert np.array_equal(sc[[0, 2, 3, 4]], c["scalars"][[0, 2, 3, 4]])


def test_percentile_index_matches_numpy():
    rng = np.random.default_rng(0)
    for n in list(range(2, 200)) + [255, 256, 1000, 1024, 2048, 4096, 8192, 16384, 32768, 65536]:
        x = rng.normal(size=n).astype(np.float32)
        for q in (40., 100, 0, 50, 99.9, 12.5):
            assert tier_p.percentile_f32(x, q) == np.percentile(x, q)


def test_audio_restatement_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "tier_p_audio.npz"))
    st = tier_p.InterpState()
    assert np.array_equal(st.h, g["h"])
    assert st.n_tap == 33 and abs(st.h[16] - 0.249914246020029) < 1e-15 and abs(st.h.sum() - 1) < 1e-15
    hist_c = np.zeros(32)
    for k in range(g["x"].shape[0]):
        buf, out = tier_p.play_buffer(g["x"][k], st, int(g["volume"][k]), float(g["balance"][k]))
        assert np.array_equal(out, g["out"][k])
        mono, out_c = c_oracle.play_buffer(g["x"][k], hist_c, st.h, int(g["volume"][k]), float(g["balance"][k]))
        assert np.abs(out_c.astype(int) - g["out"][k].astype(int)).max() <= 1    # summation order unpinned
        assert np.abs(mono - buf).max() <= 1e-9 * max(1.0, np.abs(buf).max())
    samples, rssi, ovf, seq = tier_p.snd_ingest(g["snd_msg"].tobytes())
    assert np.array_equal(samples, g["snd_samples"]) and ovf and seq == 1234 and abs(rssi - (-40.0)) < 1e-9


def test_ingest_formats():
    body = np.arange(1024, dtype=np.uint8)
    msg = b"W/F" + b"\x00" + (1).to_bytes(4, "little") + (2).to_bytes(4, "little") + (3).to_bytes(4, "little") + body.tobytes()
    x = tier_p.wf_ingest(msg)
    assert x.dtype == np.float32 and x.shape == (1024,) and np.array_equal(x, body.astype(np.float32))
    iq = np.array([1, -2, 300, -32768], dtype=">i2").tobytes()
    msg = b"SND" + bytes([0]) + (9).to_bytes(4, "little") + (1000).to_bytes(2, "big") + bytes(10) + iq
    cs, rssi, gps, seq = tier_p.iq_ingest(msg)
    assert np.array_equal(cs, np.array([1 - 2j, 300 - 32768j], np.complex64)) and seq == 9
    pal = tier_p.cutesdr_palette()
    assert pal.shape == (255, 3) and tuple(pal[0]) == (0, 0, 0) and tuple(pal[254])[:2] == (255, 0)


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_restatement_matches_imported_reference_live():
    m = ref_import.load()
    rng = np.random.default_rng(5)
    for trial in range(200):
        W = int(rng.choice([256, 1024, 2048]))
        n = int(rng.integers(1, 30))
        lines = np.clip(rng.normal(100, 10, (n, W)), 0, 255).astype(np.uint8)
        wf = m.kiwi_waterfall.__new__(m.kiwi_waterfall)
        wf.zoom, wf.wf_auto_scaling = int(rng.integers(0, 15)), True
        wf.delta_low_db, wf.delta_high_db = int(rng.integers(-9, 9)), int(rng.integers(-9, 9))
        wf.dynamic_range = wf.MIN_DYN_RANGE
        wf.spectrum = np.mean([l.astype(np.float32) for l in lines], axis=0) if n > 1 else lines[0].astype(np.float32)
        wf.spectrum_db2col()
        st = tier_p.ColourState()
        st.zoom, st.delta_low_db, st.delta_high_db = wf.zoom, wf.delta_low_db, wf.delta_high_db
        _, col, _ = tier_p.waterfall_line(lines, st)
        assert np.array_equal(col, wf.wf_color)
    f = m.filtering(6000, 48000)
    assert np.array_equal(f.h, tier_p.fir_design(6000, 48000))


def test_markstein_division_equals_ieee():
    """The kernel's division by loop-invariant divisors (csrc/wf_kernels.cu div_rn) is IEEE-exact:
    every byte sum / n_avg pair the waterfall can produce, and random colour quotients
    (w - low) / den in the ranges spectrum_db2col produces (utils_supersdr.py:793-809)."""
    for n in range(1, 101):
        a = np.arange(0, 255 * n + 1, dtype=np.float32)
        assert c_oracle.markstein_mismatches(a, n) == 0
    rng = np.random.default_rng(5)
    for _ in range(200):
        den = np.float32(rng.uniform(1.0, 400.0))
        if (den.view(np.uint32) & 0x7fffff) == 0x7fffff:
            continue
        a = rng.uniform(-400.0, 400.0, 20000).astype(np.float32)
        a[:100] = (rng.integers(-2000, 2000, 100) / np.float32(10.0)).astype(np.float32)
        a[100] = 0.0
        assert c_oracle.markstein_mismatches(a, den) == 0
