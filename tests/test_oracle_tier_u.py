"""CPU: the builder-defined Tier-U oracles against each other and against their committed fixtures."""
import os

import numpy as np

from oracle import tier_u, c_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_c_fft_is_an_fft():
    rng = np.random.default_rng(0)
    for N in (64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536):
        x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
        _, spec = c_oracle.wf_frame_bytes(x, window=False, want_spectrum=True)
        ref = np.fft.fft(x.astype(np.complex128))
        err = np.sqrt(np.mean(np.abs(spec - ref) ** 2)) / np.sqrt(np.mean(np.abs(ref) ** 2))
        assert err < 5e-7, (N, err)


def test_c_bytes_vs_float64_boundary_aware():
    for N in (256, 1024, 4096, 16384, 65536):
        x = tier_u.synth_iq(N, seed=N)[0]
        by = c_oracle.wf_frame_bytes(x)
        mism, unexplained = tier_u.compare_bytes_boundary_aware(by, x)
        assert unexplained == 0 and mism <= max(2, N // 500)
    # silence -> byte 0 everywhere; full-scale tone at an exact bin -> 245 (= -10 dBm + 255)
    assert c_oracle.wf_frame_bytes(np.zeros(1024, np.complex64)).max() == 0
    n = np.arange(1024)
    tone = (32768 * np.exp(2j * np.pi * 100 * n / 1024)).astype(np.complex64)
    by = c_oracle.wf_frame_bytes(tone)
    assert by[512 + 100] == 245 and by.argmax() == 612


def test_golden_waterfall():
    g = np.load(os.path.join(GOLD, "tier_u_waterfall.npz"))
    assert np.array_equal(c_oracle.wf_frame_bytes(g["iq1"][0]), g["bytes1"])
    r = c_oracle.wf_rows(g["iq1"][None], zoom=0)
    assert np.array_equal(r["pixels"], g["pixels1"]) and np.array_equal(r["spectrum"], g["spectrum1"])
    r2 = c_oracle.wf_rows(g["iq2"], zoom=3, threads=2)
    assert np.array_equal(r2["pixels"], g["pixels2"]) and np.array_equal(r2["sums"], g["sums2"])
    assert np.array_equal(r2["scalars"], g["scalars2"])


def test_plan_and_tables():
    # one first pass of radix N / 32^k, then k radix-32 passes
    assert c_oracle.fft_plan(1024) == [32, 32] and c_oracle.fft_plan(16384) == [16, 32, 32]
    assert c_oracle.fft_plan(512) == [16, 32] and c_oracle.fft_plan(256) == [8, 32]
    assert c_oracle.fft_plan(2048) == [2, 32, 32] and c_oracle.fft_plan(8192) == [8, 32, 32]
    # larger than one SM's shared memory: a front pass over HBM, then the 16384-point plan
    assert c_oracle.fft_plan(32768) == [2, 16, 32, 32] and c_oracle.fft_plan(65536) == [4, 16, 32, 32]
    # the window table is the first half of the periodic Hann window; the second half is 1 - w
    w = c_oracle.window_table(1024).astype(np.float64)
    assert w.size == 512 and w[0] == 0 and abs(w[256] - 0.5) < 1e-7
    assert np.allclose(w, tier_u.hann(1024)[:512], atol=6e-8) and np.allclose(1 - w, tier_u.hann(1024)[512:], atol=6e-8)
    t = c_oracle.thresholds(1024, -10.0)
    assert t[0] == 0 and np.all(np.diff(t[1:]) > 0)
    # threshold k sits half a dB below byte k: 10 log10(T[k]/ref) - 10 + 255 == k - 0.5
    ref = (1024 * 32768.0 * 0.5) ** 2
    k = np.arange(1, 256)
    assert np.allclose(10 * np.log10(t[1:].astype(np.float64) / ref) - 10 + 255, k - 0.5, atol=1e-5)


def test_demod_golden_and_properties():
    g = np.load(os.path.join(GOLD, "tier_u_demod.npz"))
    for mode in tier_u.MODES:
        p = tier_u.DemodParams(mode, decay=1000 if mode == "cw" else 4000, hang=(mode == "cw"))
        st = tier_u.DemodState()
        pcm, rssi = tier_u.demod(g["iq_" + mode], p, st)
        assert np.allclose(pcm, g["pcm_" + mode], rtol=1e-5, atol=1e-2)
        assert np.allclose(rssi, g["rssi_" + mode], atol=1e-3)
        # streaming invariance: frame-by-frame == one shot
        st2 = tier_u.DemodState()
        parts = [tier_u.demod(g["iq_" + mode][i:i + 512], p, st2)[0] for i in range(0, 4096, 512)]
        assert np.array_equal(np.concatenate(parts), pcm)
    # USB passes +1 kHz and rejects -1 kHz
    x = tier_u.synth_demod_iq("usb", 512 * 24, seed=1)
    pcm, _ = tier_u.demod(x, tier_u.DemodParams("usb"), tier_u.DemodState())
    F = np.abs(np.fft.rfft(pcm[-4096:] * np.hanning(4096)))
    assert abs(np.fft.rfftfreq(4096, 1 / 12000)[F.argmax()] - 1000) < 5


def test_envelope_scan_form_equals_recurrence():
    rng = np.random.default_rng(0)
    mag = np.abs(rng.standard_normal(512)) * 100
    ring = np.abs(rng.standard_normal(11)) * 150
    c2 = np.log2(np.e) / (12000 * 0.4)
    for hang in (False, True):
        a = tier_u.demod_block_envelope(mag, hang, 220.0, ring, np.exp2(-c2))
        b = tier_u.demod_block_envelope_scan(mag, hang, 220.0, ring, c2)
        assert np.abs(a - b).max() / a.max() < 1e-13
