"""CPU: numerical model of the tcgen05 demodulator engine's split-precision FIR: the mixed signal goes
to the tensor core as hi = cvt.rna.tf32(x) plus lo = bfloat16(x - hi), the taps as B_hi = tf32(h), B_lo = h - B_hi
(truncated to TF32 by the tensor core) and bfloat16(h); D = hi [B_hi | B_lo] + lo bf16(h), accumulated in float32.
The model restates that arithmetic in numpy and checks the error budget against the float64 FIR for the signal
classes the demodulator sees -- far inside the 1e-5 relative RMS tolerance the GPU tests hold both engines to."""
import numpy as np
import pytest

from oracle import tier_u


def _tf32_rna(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    return ((u + 0x1000) & 0xFFFFE000).astype(np.uint32).view(np.float32)


def _tf32_trunc(x):
    return (np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


def _bf16_rn(x):
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u & 0xFFFF0000).astype(np.uint32).view(np.float32)


def _fir_split(z, h):
    """z float32[n] (one of re / im), h float32[127] -> float32[n - 126]: the engine's three products, float32 accumulation."""
    hi = _tf32_rna(z)
    lo = _bf16_rn(z - hi)
    assert np.array_equal(hi.astype(np.float64) + (z - hi).astype(np.float64), z.astype(np.float64))   # the split is exact
    b_hi = _tf32_rna(h)
    b_lo = _tf32_trunc(h - b_hi)
    b16 = _bf16_rn(h)
    win = np.lib.stride_tricks.sliding_window_view

    def dot(a, b):                                           # one accumulator: float32 rounding of the column sums
        return (win(a, 127)[:, ::-1].astype(np.float64) * b.astype(np.float64)).sum(1).astype(np.float32)
    main = (dot(hi, b_hi) + dot(lo, b16)).astype(np.float32)   # TMEM columns 0..31
    return main + dot(hi, b_lo)                               # + columns 32..63 after the read-back


@pytest.mark.parametrize("kind", ["usb_tone", "full_scale", "weak_plus_strong", "tiny", "clipping"])
def test_split_precision_fir_error_budget(kind):
    n = 4096
    t = np.arange(n)
    rng = np.random.default_rng(5)
    if kind == "usb_tone":
        z = tier_u.synth_demod_iq("usb", n, seed=2).real
    elif kind == "full_scale":
        z = 32767.0 * np.cos(2 * np.pi * 700.0 * t / 12000)
    elif kind == "clipping":                                # |I + jQ| mixed onto one axis: sqrt(2) x full scale
        z = 46340.0 * np.sign(np.cos(2 * np.pi * 300.0 * t / 12000))
    elif kind == "weak_plus_strong":                       # a weak in-band tone next to a 70 dB stronger out-of-band one
        z = 30000.0 * np.cos(2 * np.pi * 4900.0 * t / 12000) + 10.0 * np.cos(2 * np.pi * 500.0 * t / 12000)
    else:
        z = 1e-2 * rng.standard_normal(n)                   # far below one ADC count
    z = z.astype(np.float32)
    h = tier_u.demod_taps(300.0, 2700.0).astype(np.float32)
    ref = np.convolve(z.astype(np.float64), h.astype(np.float64), "valid")
    got = _fir_split(z, h).astype(np.float64)
    # relative to the level the filter sees, like any float32 FIR: the bfloat16 roundings of the correction term leave
    # ~3e-7 of the input level (-130 dBc; the int16 input's own quantisation noise is at -101 dBFS)
    err = np.sqrt(np.mean((got - ref) ** 2))
    assert err / np.sqrt(np.mean(z.astype(np.float64) ** 2)) < 5e-7
    if kind != "weak_plus_strong":                           # in-band signals: the output is that good, too
        assert err / np.sqrt(np.mean(ref ** 2)) < 2e-6


def test_tiny_taps_keep_their_relative_accuracy():
    """Why the correction operands are bfloat16 and not (scaled) float16: a +-6 kHz filter at 12 kHz is a unit tap plus
    taps of ~1e-17; in the start-up transient of a stream only those contribute, and the NBFM detector takes the PHASE
    of that output.  bfloat16 has float32's exponent range, float16 would flush the taps to zero and lose the lo part
    (2^-12 relative, 20 x the tolerance over the first 63 samples)."""
    h = tier_u.demod_taps(-6000.0, 6000.0).astype(np.float32)
    assert abs(h[63] - 1.0) < 1e-6 and 0 < np.abs(h[1:63]).max() < 1e-12
    z = (1000.0 * np.cos(2 * np.pi * 1000.0 * np.arange(400) / 12000) + 3.0).astype(np.float32)
    zp = np.concatenate([np.zeros(126, np.float32), z])     # a stream starts from zero history
    ref = np.convolve(zp.astype(np.float64), h.astype(np.float64), "valid")[1:60]
    got = _fir_split(zp, h).astype(np.float64)[1:60]
    assert np.all(np.abs(got - ref) <= 2e-6 * np.abs(ref))
    h16 = h.astype(np.float16).astype(np.float32)
    assert np.count_nonzero(h16[1:63]) == 0


def test_two_term_split_would_miss_the_tolerance():
    """Why the lo part is needed: hi [B_hi | B_lo] alone leaves ~2^-12 relative error per sample."""
    n = 4096
    z = (32767.0 * np.cos(2 * np.pi * 700.0 * np.arange(n) / 12000)).astype(np.float32)
    h = tier_u.demod_taps(300.0, 2700.0).astype(np.float32)
    ref = np.convolve(z.astype(np.float64), h.astype(np.float64), "valid")
    two = np.convolve(_tf32_rna(z).astype(np.float64), h.astype(np.float64), "valid")
    assert np.sqrt(np.mean((two - ref) ** 2)) / np.sqrt(np.mean(ref ** 2)) > 1e-5
