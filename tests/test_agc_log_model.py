"""CPU restatement of the AGC arithmetic the demodulator kernels evaluate (supersdr_b200/csrc/demod_post.cuh, round 2:
envelope and gain in the log2-of-POWER domain, float32, lane-local prefix + cross-lane scan) against the oracle's
sequential float64 envelope (oracle/tier_u.demod_block_envelope).  No GPU: this pins the algebra and
the float32 error budget (the GPU parity tests then cover the real kernels at 1e-5 RMS).  Parameter model:
utils_supersdr.py:936-945, 1022-1029 (thresh / slope / decay / hang)."""
import numpy as np
import pytest

from oracle import tier_u as U

f32 = np.float32


def _lg2(x):
    with np.errstate(divide="ignore"):
        return np.log2(x.astype(np.float64)).astype(f32)


def _ex2(x):
    return np.exp2(x.astype(np.float64)).astype(f32)


def _fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)


def frame_gain(pw, hang, e_in, ring, c2, knee2, slope_m1):
    """One 512-sample frame: pw = |z|^2 (float32), 32 lanes x 16 samples.  Returns (gain[512], e_in of the next frame)."""
    pw = pw.astype(f32).reshape(32, 16)
    if hang:
        hb = f32(ring.max())
        run = np.maximum.accumulate(pw, axis=1)
        incl = np.maximum.accumulate(run[:, 15])
        excl = np.concatenate([[f32(0)], incl[:-1]]).astype(f32)
        q = np.maximum(run, np.maximum(excl, hb * hb)[:, None])
    else:
        q = pw
    c2d = f32(2) * f32(c2)
    r = np.broadcast_to(np.arange(16, dtype=f32)[None, :], (32, 16))
    W = np.maximum.accumulate(_fma(r, c2d, _lg2(q)), axis=1)
    base = (np.arange(32) * 16).astype(f32) * c2d
    pre = np.maximum.accumulate((W[:, 15] + base).astype(f32))
    pre = np.concatenate([[f32(-np.inf)], pre[:-1]]).astype(f32)
    seed = _fma(f32(2), _lg2(np.array([e_in], f32)), -c2d)[0]
    pre = (np.maximum(pre, seed) - base).astype(f32)
    knee, gs = f32(2) * (f32(knee2) + f32(15)), f32(0.5) * f32(slope_m1)
    gc = _fma(f32(-15), f32(slope_m1), f32(-1))
    e2 = _fma(-r, c2d, np.maximum(W, pre[:, None]))
    g = _ex2(_fma(np.maximum(e2, knee), gs, gc))
    return g.reshape(-1), _ex2(np.array([f32(0.5) * e2[31, 15]], f32))[0]


@pytest.mark.parametrize("trial", range(10))
def test_log_power_agc_matches_sequential_envelope(trial):
    rng = np.random.default_rng(100 + trial)
    hang = trial % 2
    decay = [1.0, 50, 400, 1000, 8000][trial % 5]
    thresh = -100 + 10 * (trial % 7)
    slope = [6, 0, 30][trial % 3]
    c2 = np.log2(np.e) / (12000 * decay / 1000)
    knee2 = (thresh - U.AGC_FS_DBM) / 20 * np.log2(10)
    s = slope / 100
    nfr = 24
    mag = np.abs(rng.normal(size=nfr * 512)) * 10 ** rng.uniform(0, 4.3) * np.repeat(10 ** rng.uniform(-2, 0, nfr), 512)
    if trial % 4 == 0:
        mag[512 * 3:512 * 9] = 0                            # silence in the middle of a stream
    if trial == 9:
        mag[:] = 0                                          # all silence: gain stays at the knee, nothing becomes NaN
    mag = mag.astype(f32).astype(np.float64)
    e_o, ring = 0.0, np.zeros(U.HANG_BLOCKS)
    e_m, ring_m = f32(0), np.zeros(U.HANG_BLOCKS, f32)
    worst = 0.0
    for b in range(nfr):
        mb = mag[b * 512:(b + 1) * 512]
        e = U.demod_block_envelope(mb, hang, e_o, ring, np.exp2(-c2))
        with np.errstate(divide="ignore"):
            m2 = np.log2(e / U.FS)
        g_ref = U.AGC_OUT * np.exp2(np.maximum(m2, knee2) * (s - 1))
        e_o = e[-1]
        ring[b % U.HANG_BLOCKS] = mb.max()
        g, e_m = frame_gain((mb.astype(f32) ** 2).astype(f32), hang, e_m, ring_m, c2, knee2, s - 1)
        ring_m[b % U.HANG_BLOCKS] = np.sqrt(f32(mb.max()) ** 2)
        assert np.all(np.isfinite(g))
        worst = max(worst, np.abs(g / g_ref - 1).max())
    assert worst < 1e-5, worst
