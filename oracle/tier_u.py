"""TEST INFRASTRUCTURE ONLY -- Tier-U oracle (PARITY UNPINNED, builder-defined spec).

The reference client never computes an FFT, a log-magnitude or a demodulated sample: that
arithmetic lives in the KiwiSDR server firmware (jks-prv/Beagle_SDR_GPS -- not vendored, not a
declared dependency, no pinned version).  This module is the float64
statement of the spec frozen in DESIGN.md section 4, anchored only on the wire conventions the
reference pins:

* waterfall byte = dBm + 255                        utils_supersdr.py:789
* 1024 bins, bin 0 = lowest frequency               utils_supersdr.py:596,772-774
* 12 kHz IQ/audio, 512-sample frames                utils_supersdr.py:905-909
* mode pass-bands                                   utils_supersdr.py:42-50,859-873; kiwi/client.py:217-253
* AGC parameter names / units / defaults            utils_supersdr.py:936-945,1023
* IQ samples = big-endian int16 counts              kiwi/client.py:443-454

It is the ACCURACY oracle (float64); the bit-exact float32 statement of the waterfall stage is
oracle/c/ssdr_oracle.c.  The demodulator is compared at the north_star tolerance (1e-5 RMS).
"""
import numpy as np

FS = 32768.0                 # IQ full scale (int16 counts)
WF_CAL_DB = -10.0            # dBFS -> dBm: a full-scale tone reads -10 dBm (byte 245)
KIWI_RATE = 12000
FRAME = 512                  # KIWI_SAMPLES_PER_FRAME, utils_supersdr.py:909

# ---------------------------------------------------------------------------------------------
# waterfall: IQ frame -> Kiwi byte line
# ---------------------------------------------------------------------------------------------


def hann(N):
    """Periodic Hann, w[n] = 0.5 - 0.5 cos(2 pi n / N)."""
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(N) / N)


def wf_frame_db(iq, window=True, cal_db=WF_CAL_DB):
    """complex[N] -> float64 'byte units' (dBm + 255) per bin, fftshifted, before rounding."""
    x = np.asarray(iq, dtype=np.complex128)
    N = x.size
    if window:
        x = x * hann(N)
    X = np.fft.fftshift(np.fft.fft(x))
    P = X.real ** 2 + X.imag ** 2
    ref = (N * FS * 0.5) ** 2
    with np.errstate(divide="ignore"):
        db = 10.0 * np.log10(P / ref) + cal_db
    return db + 255.0, np.sqrt(P)


def wf_frame_bytes(iq, window=True, cal_db=WF_CAL_DB):
    v, _ = wf_frame_db(iq, window, cal_db)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def fft_error_bound(iq, window=True):
    """A generous bound on |X_fp32 - X_exact| per bin for a float32 FFT of this frame:
    c * eps32 * log2(N) * ||x w||_2 (random-walk growth is ~sqrt(log2 N); we allow the linear
    worst case with c = 2)."""
    x = np.asarray(iq, dtype=np.complex128)
    N = x.size
    if window:
        x = x * hann(N)
    return 2.0 * np.finfo(np.float32).eps * np.log2(N) * np.sqrt(np.sum(np.abs(x) ** 2))


def compare_bytes_boundary_aware(got_u8, iq, window=True, cal_db=WF_CAL_DB):
    """Boundary-aware comparator: a float32 result may differ from the float64
    rounding only by +-1 and only where the float64 value lies within the float32 error band of a
    rounding boundary.  Returns (n_mismatch, n_unexplained)."""
    v, amp = wf_frame_db(iq, window, cal_db)
    exact = np.clip(np.rint(v), 0, 255).astype(np.int64)
    got = np.asarray(got_u8).astype(np.int64)
    err = fft_error_bound(iq, window)
    with np.errstate(divide="ignore", invalid="ignore"):
        band = 20.0 * np.log10(1.0 + err / np.maximum(amp, 1e-300))     # dB the value may move
    band = np.where(np.isfinite(band), band, 1e9) + 1e-6
    frac = np.abs(v - np.floor(v) - 0.5)                                   # distance to a boundary
    diff = got - exact
    mism = diff != 0
    # values clipped at 0 are fully explained when the exact value is below the first boundary
    explained = (np.abs(diff) <= 1) & (frac <= band)
    explained |= (v < 0.5 + band) & (got <= 1) & (exact <= 1)
    return int(mism.sum()), int((mism & ~explained).sum())


# ---------------------------------------------------------------------------------------------
# synthetic inputs
# ---------------------------------------------------------------------------------------------

def synth_iq(N, seed=1234, tones=((0.5, 100.0), (0.05, 300.5), (0.005, 777.25)), sigma=1e-3,
             frames=1, quantise=False):
    """Tones (amplitude in FS, frequency in FFT bins of an N-point frame) + complex AWGN.
    Returns complex64[frames, N] in int16-count units (optionally rounded to integer counts as
    a real Kiwi IQ stream would be)."""
    rng = np.random.default_rng(seed)
    n = np.arange(frames * N, dtype=np.float64)
    x = np.zeros(frames * N, dtype=np.complex128)
    for amp, b in tones:
        ph = rng.uniform(0, 2 * np.pi)
        x += amp * np.exp(1j * (2 * np.pi * b * n / N + ph))
    x += sigma * (rng.standard_normal(frames * N) + 1j * rng.standard_normal(frames * N)) / np.sqrt(2)
    x *= FS
    if quantise:
        x = np.rint(x.real) + 1j * np.rint(x.imag)
    return x.astype(np.complex64).reshape(frames, N)


def synth_batch(B, n, N, seed=1234, quantise=False):
    """complex64[B, n, N]; channel ch uses seed+ch and random tone bins."""
    out = np.empty((B, n, N), np.complex64)
    for ch in range(B):
        rng = np.random.default_rng(seed + ch)
        bins = rng.uniform(0, N, 3)
        tones = ((0.5, bins[0]), (0.05, bins[1]), (0.005, bins[2]))
        out[ch] = synth_iq(N, seed + ch, tones, frames=n, quantise=quantise)
    return out


# ---------------------------------------------------------------------------------------------
# demodulator: IQ @12 kHz -> PCM @12 kHz     (ABSENT from the reference)
# ---------------------------------------------------------------------------------------------

MODES = {"am": 0, "usb": 1, "lsb": 2, "cw": 3, "nbfm": 4}
FIR_TAPS = 127               # fixed odd length of the real low-pass prototype
AGC_FS_DBM = -10.0           # an IQ tone of amplitude FS reads -10 dBm (same cal as the waterfall)
AGC_OUT = 0.5                # output peak (fraction of FS) of a full-scale input at slope 0
HANG_BLOCKS = 11             # hang=1 holds a peak for the rest of its block + 11 blocks (~0.5 s)
AM_DC_TAU = 0.1              # s, AM carrier (DC) tracker time constant


def default_passband(mode):
    """SuperSDR pass-bands, utils_supersdr.py:42-50,859-873; NBFM from kiwi/client.py:237-239."""
    return {"usb": (30, 3000), "lsb": (-3000, -30), "am": (-6000, 6000), "cw": (400, 800),
            "nbfm": (-6000, 6000)}[mode]


class DemodParams:
    """Per-channel parameter block.  AGC names/defaults: utils_supersdr.py:936-945."""

    def __init__(self, mode="usb", lc=None, hc=None, f_off=0.0, on=True, hang=False, thresh=-80,
                 slope=0, decay=4000, gain=50):
        self.mode = mode
        dlc, dhc = default_passband(mode)
        self.lc = dlc if lc is None else lc
        self.hc = dhc if hc is None else hc
        self.f_off = f_off
        self.on, self.hang, self.thresh, self.slope, self.decay, self.gain = on, hang, thresh, slope, decay, gain


def demod_taps(lc, hc, fs=KIWI_RATE, T=FIR_TAPS):
    """Real low-pass prototype of the pass-band filter: cut-off (hc-lc)/2, Blackman-windowed sinc,
    unity DC gain -- the recipe of ``filtering`` (utils_supersdr.py:333-344) at a fixed length."""
    fl = (hc - lc) / 2.0
    h = np.sinc(2.0 * fl / fs * (np.arange(T) - (T - 1) / 2.0)) * np.blackman(T)
    h = h / np.sum(h)
    # the Blackman window is exactly 0 at both ends (0.42 - 0.5 + 0.08); float64
    # leaves +-1.4e-17 there, which would make the first output sample of a stream pure rounding
    # noise (and its NBFM phase arbitrary).  The spec pins the two end taps to exactly 0.
    h[0] = h[-1] = 0.0
    return h


def phase_inc(f_hz, fs=KIWI_RATE):
    """32-bit phase-accumulator increment."""
    return int(np.rint(f_hz / fs * 4294967296.0)) % 4294967296


class DemodState:
    def __init__(self, T=FIR_TAPS):
        self.ph1 = 0
        self.ph2 = 0
        self.hist = np.zeros(T - 1, np.complex128)
        self.e_in = 0.0
        self.ring = np.zeros(HANG_BLOCKS)
        self.blk = 0
        self.dc = 0.0
        self.zprev = 0.0 + 0.0j


def _nco(ph0, inc, n):
    ph = (ph0 + inc * np.arange(n, dtype=np.uint64)) % 4294967296
    # signed phase in (-pi, pi]: identical value of exp(), friendlier to float32 implementations
    sph = ph.astype(np.int64)
    sph = np.where(sph >= 2147483648, sph - 4294967296, sph)
    return 2.0 * np.pi * sph.astype(np.float64) / 4294967296.0, int((ph0 + inc * n) % 4294967296)


def demod_block_envelope(mag, hang, e_in, ring, alpha):
    """Peak envelope of one 512-block, sequential definition:
    hm[k] = max(mag[0..k], max(ring)) if hang else mag[k];  e[k] = max(hm[k], alpha * e[k-1])."""
    hm = np.maximum(np.maximum.accumulate(mag), ring.max()) if hang else mag
    e = np.empty_like(mag)
    prev = e_in
    for k in range(mag.size):
        prev = max(hm[k], alpha * prev)
        e[k] = prev
    return e


def demod_block_envelope_scan(mag, hang, e_in, ring, c2):
    """The same envelope in scan form (what the GPU evaluates): e[k] = 2^(-k c2) *
    max(2^(-c2) e_in, max_{i<=k} hm[i] 2^(i c2)).  Used in tests to show both forms agree."""
    k = np.arange(mag.size, dtype=np.float64)
    hm = np.maximum(np.maximum.accumulate(mag), ring.max()) if hang else mag
    u = hm * np.exp2(k * c2)
    M = np.maximum(np.maximum.accumulate(u), e_in * np.exp2(-c2))
    return M * np.exp2(-k * c2)


def demod(iq, p, st, fs=KIWI_RATE):
    """One channel, len(iq) a multiple of 512.  Returns (pcm float64, rssi dBm per block).
    ``st`` (DemodState) is updated so that consecutive calls stream."""
    x = np.asarray(iq, dtype=np.complex128)
    n = x.size
    assert n % FRAME == 0
    mode = MODES[p.mode]
    fc = (p.lc + p.hc) / 2.0
    inc1, inc2 = phase_inc(p.f_off + fc, fs), phase_inc(fc, fs)
    h = demod_taps(p.lc, p.hc, fs).astype(np.float32).astype(np.float64)   # device taps are float32
    th1, st.ph1 = _nco(st.ph1, inc1, n)
    z1 = x * np.exp(-1j * th1)
    zz = np.concatenate([st.hist, z1])
    st.hist = zz[-(FIR_TAPS - 1):].copy()
    z2 = np.convolve(zz, h, mode="valid")                    # z2[k] = sum_t h[t] z1[k - t]
    th2, st.ph2 = _nco(st.ph2, inc2, n)
    mag = np.abs(z2)
    tau = p.decay / 1000.0
    c2 = np.log2(np.e) / (fs * tau)
    alpha = np.exp2(-c2)
    knee2 = (p.thresh - AGC_FS_DBM) / 20.0 * np.log2(10.0)
    s = p.slope / 100.0
    beta = 1.0 / (fs * AM_DC_TAU)
    pcm = np.empty(n)
    rssi = np.empty(n // FRAME)
    for b in range(n // FRAME):
        sl = slice(b * FRAME, (b + 1) * FRAME)
        zb, mb = z2[sl], mag[sl]
        rssi[b] = 10.0 * np.log10(max(np.mean(mb ** 2), 1e-30) / FS ** 2) + AGC_FS_DBM
        if mode == 4:                                        # NBFM: quadrature detector, no AGC
            prev = np.concatenate([[st.zprev], zb[:-1]])
            prod = zb * np.conj(prev)
            # a zero product (first sample of a stream, or silence) demodulates to 0, not +-pi
            pcm[sl] = np.where(prod == 0, 0.0, np.angle(prod)) * (32767.0 / np.pi)
            st.zprev = zb[-1]
        else:
            if mode == 0:                                    # AM: envelope minus tracked carrier
                a = np.empty(FRAME)
                dc = st.dc
                for k in range(FRAME):
                    dc = dc + beta * (mb[k] - dc)
                    a[k] = mb[k] - dc
                st.dc = dc
            else:                                            # USB / LSB / CW: product detector
                a = (zb * np.exp(1j * th2[sl])).real
            if p.on:
                e = demod_block_envelope(mb, p.hang, st.e_in, st.ring, alpha)
                with np.errstate(divide="ignore"):
                    m2 = np.log2(e / FS)
                gain = AGC_OUT * np.exp2(np.maximum(m2, knee2) * (s - 1.0))
                st.e_in = e[-1]
            else:
                gain = 10.0 ** (p.gain / 20.0)
            pcm[sl] = a * gain
            st.zprev = zb[-1]
        st.ring[st.blk % HANG_BLOCKS] = mb.max()
        st.blk += 1
    return pcm, rssi


def pcm_to_i16(pcm):
    return np.clip(np.rint(pcm), -32768, 32767).astype(np.int16)


def synth_demod_iq(mode, n, seed=0, fs=KIWI_RATE, sigma=1e-3, level=0.1):
    """Per-mode test signal, complex64 int16-count units."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / fs
    noise = sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(2)
    if mode in ("usb", "lsb"):
        # in-band tone at +-1 kHz and an opposite-sideband tone that must be rejected
        sgn = 1.0 if mode == "usb" else -1.0
        x = level * np.exp(2j * np.pi * sgn * 1000.0 * t) + level * np.exp(-2j * np.pi * sgn * 1000.0 * t + 1.0j)
    elif mode == "cw":
        key = (np.floor(t / 0.06) % 2 == 0).astype(float)     # ~10 Hz keying
        x = level * key * np.exp(2j * np.pi * 600.0 * t)
    elif mode == "am":
        x = level * (1.0 + 0.3 * np.sin(2 * np.pi * 440.0 * t))
    else:                                                     # nbfm: 1 kHz tone, 3 kHz deviation
        x = level * np.exp(1j * (3000.0 / 1000.0) * np.sin(2 * np.pi * 1000.0 * t))
    return ((x + noise) * FS).astype(np.complex64)
