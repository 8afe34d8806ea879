"""CPU restatement of the NBFM detector's arctangent (supersdr_b200/csrc/demod_post.cuh: atan2_fast2_scaled, round 2):
octant reduction t = min / max, odd polynomial of degree 15 in t with the output scale K = 32767 / pi folded into its
coefficients, octant fix-ups with K pi / 2 and K pi, sign of y.  float32 arithmetic with fused multiply-adds, as the
kernel evaluates it (two samples per packed FFMA2; the lanes are independent, so one lane is restated).  Pins the error
DESIGN.md 4.5 states: <= 1.5e-7 rad of polynomial error plus the roundings of the reduction.  The detector itself
(phase of z[n] conj(z[n-1])) is checked against the float64 oracle by the GPU parity tests."""
import numpy as np

f32 = np.float32
K = f32(32767.0) / f32(3.14159265358979)
COEF = [-0.00455979211255908, 0.023780519142746925, -0.05882975459098816, 0.09868865460157394, -0.14003290235996246,
        0.19966961443424225, -0.3333181142807007, 0.9999998807907104]


def _fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def atan2_scaled_model(y, x):
    y, x = y.astype(f32), x.astype(f32)
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (mn * (f32(1) / mx).astype(f32)).astype(f32)           # rcp.approx + multiply
    q = (t * t).astype(f32)
    ck = [(f32(c) * K).astype(f32) for c in COEF]
    p = np.full_like(t, ck[0])
    for c in ck[1:]:
        p = _fma(p, q, np.full_like(t, c))
    r = (p * t).astype(f32)
    r = np.where(ay > ax, (f32(1.57079637) * K).astype(f32) - r, r).astype(f32)
    r = np.where(x < 0, (f32(3.14159274) * K).astype(f32) - r, r).astype(f32)
    return np.copysign(r, y)


def test_scaled_atan2_error_bound():
    rng = np.random.default_rng(7)
    n = 400000
    mag = 10.0 ** rng.uniform(-3, 9, n)
    ang = rng.uniform(-np.pi, np.pi, n)
    y, x = (mag * np.sin(ang)).astype(f32), (mag * np.cos(ang)).astype(f32)
    ok = (x != 0) | (y != 0)
    got = atan2_scaled_model(y[ok], x[ok]).astype(np.float64) / float(K)
    ref = np.arctan2(y[ok].astype(np.float64), x[ok].astype(np.float64))
    err = np.abs(got - ref)
    err = np.minimum(err, 2 * np.pi - err)                         # +pi and -pi are the same phase
    assert err.max() < 6e-7, err.max()                             # 1.5e-7 polynomial + float32 roundings at |phase| ~ pi (ulp 2.4e-7)
    assert np.sqrt(np.mean(err ** 2)) < 1.5e-7


def test_scaled_atan2_axes_and_octant_edges():
    v = np.array([1.0, 3.0, 1e-3, 2.5e4], f32)
    for s in v:
        z = f32(0)
        cases = [(z, s, 0.0), (s, z, np.pi / 2), (z, -s, np.pi), (-s, z, -np.pi / 2), (s, s, np.pi / 4), (s, -s, 3 * np.pi / 4),
                 (-s, -s, -3 * np.pi / 4), (-s, s, -np.pi / 4)]
        for y, x, want in cases:
            got = float(atan2_scaled_model(np.array([y], f32), np.array([x], f32))[0]) / float(K)
            d = abs(got - want)
            assert min(d, 2 * np.pi - d) < 5e-7, (y, x, got, want)
