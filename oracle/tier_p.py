"""TEST INFRASTRUCTURE ONLY -- Tier-P oracle (PINNED).

Plain-numpy restatement of the arithmetic that the reference client really performs.  Each function
cites the reference lines it follows (paths relative to /root/reference).  The restatement is
checked bit-for-bit against the unmodified reference functions (``tests/test_oracle_tier_p.py``,
via ``oracle.ref_import`` when /root/reference is present) and against the committed fixtures
``tests/golden/tier_p_*.npz`` that ``oracle/make_golden.py`` produced from them.

All waterfall arithmetic is float32 (numpy NEP-50: Python scalars are weak, so the reference's
``wf - 13 + 3*zoom`` etc. stay float32).  All audio arithmetic is float64.
"""
import struct

import numpy as np

f32 = np.float32

# ---------------------------------------------------------------------------------------------
# waterfall
# ---------------------------------------------------------------------------------------------

WF_HEADER_BYTES = 16          # "W/F" + 1 skip byte + <III (x_bin, flags|zoom, seq)   utils:782


def wf_ingest(msg):
    """utils_supersdr.py:780-784 -- strip the 16-byte header, bytes -> float32 (Kiwi byte units)."""
    if msg[0:3] != b"W/F":
        return None
    body = bytes(msg[WF_HEADER_BYTES:])
    return np.frombuffer(body, dtype=np.uint8).astype(np.float32)


def time_bin_mean(lines):
    """utils_supersdr.py:881-886 -- ``np.mean(deque of n float32[W], axis=0)``.

    The float32 row sum is exact (integers <= 100*255 < 2**24); one IEEE float32 divide by n.
    """
    lines = np.asarray(lines, dtype=np.float32)
    if lines.ndim == 1:
        return lines.copy()
    s = np.zeros(lines.shape[1], dtype=np.float32)
    for row in lines:                      # sequential row adds, as numpy's axis-0 reduction does
        s = s + row
    return s / f32(lines.shape[0])


def percentile_virtual_index(n, q_percent):
    """numpy's 'linear' percentile index for float32 data, restated.

    numpy/lib/_function_base_impl.py (numpy 2.3): ``q = true_divide(q, float32(100))`` (float32
    because the data are float32), method 'linear': ``virtual_index = (n - 1) * q`` in float32;
    ``lo = floor(vi)``; ``gamma = vi - lo``; indices at/above n-1 select the maximum.
    Returns (lo, gamma float32).
    """
    q = f32(q_percent) / f32(100)
    vi = f32(n - 1) * q
    lo = int(np.floor(vi))
    gamma = f32(vi - f32(lo))
    if lo >= n - 1:                        # numpy clips the upper neighbour index
        lo, gamma = n - 1, f32(0)
    return lo, gamma


def lerp_f32(a, b, t):
    """numpy ``_lerp`` for float32 scalars: a + (b-a)*t, switching to b - (b-a)*(1-t) for t>=0.5."""
    a, b, t = f32(a), f32(b), f32(t)
    d = f32(b - a)
    if t >= f32(0.5):
        return f32(b - f32(d * f32(f32(1) - t)))
    return f32(a + f32(d * t))


def percentile_f32(x, q_percent):
    """np.percentile(x float32, q) restated (selection + float32 lerp)."""
    x = np.asarray(x, dtype=np.float32)
    n = x.size
    lo, gamma = percentile_virtual_index(n, q_percent)
    s = np.sort(x)
    hi = min(lo + 1, n - 1)
    return lerp_f32(s[lo], s[hi], gamma)


class ColourState:
    """The attributes ``spectrum_db2col`` reads/writes on ``kiwi_waterfall`` (utils:592-603)."""
    MIN_DYN_RANGE = 40.0
    CLIP_LOWP, CLIP_HIGHP = 40.0, 100.0

    def __init__(self):
        self.low_clip_db = f32(-120)       # class defaults utils:600
        self.high_clip_db = f32(-60)
        self.dynamic_range = f32(40.0)     # utils:620
        self.delta_low_db = 0
        self.delta_high_db = 0
        self.wf_auto_scaling = True
        self.zoom = 0
        self.wf_min_db = f32(-120)
        self.wf_max_db = f32(-80)


def spectrum_db2col(spectrum, st):
    """utils_supersdr.py:787-813, float32 throughout.  Returns wf_color float32[W] in [0, 254]."""
    x = np.asarray(spectrum, dtype=np.float32)
    wf = -(f32(255) - x)                                   # :789  dBm
    wf_db = (wf - f32(13)) + f32(3 * st.zoom)              # :790
    wf_db[0] = wf_db[1]                                    # :791
    if st.wf_auto_scaling:                                 # :793-797
        st.low_clip_db = percentile_f32(wf_db, st.CLIP_LOWP)
        st.high_clip_db = f32(wf_db.max())                 # percentile 100 == max
        st.dynamic_range = f32(max(f32(st.high_clip_db - st.low_clip_db), f32(st.MIN_DYN_RANGE)))
    low = f32(f32(st.low_clip_db) + f32(st.delta_low_db))  # :800
    normal_factor_db = f32(f32(st.dynamic_range) + f32(st.delta_high_db))   # :802
    den = f32(normal_factor_db - f32(st.delta_low_db))     # :803
    c = (wf_db - low) / den
    c = np.clip(c, f32(0.0), f32(1.0))                     # :805
    st.wf_min_db = f32(low - f32(3 * st.zoom))             # :807
    st.wf_max_db = f32(f32(f32(st.low_clip_db) + normal_factor_db) - f32(3 * st.zoom))   # :808
    c = c * f32(254)                                       # :811
    c = np.clip(c, f32(0), f32(255))                       # :813
    return c.astype(np.float32)


def pixel_row(wf_color):
    """Row -> 8-bit palette indices.  The float->uint8 step happens inside pygame
    (``surfarray.make_surface``, supersdr.py:929) which is not importable here: BUILDER DEFINITION
    ``uint8(rint(c))`` (round-half-even), parity unpinned for this one cast."""
    return np.rint(np.asarray(wf_color, dtype=np.float32)).astype(np.uint8)


def waterfall_line(frames_u8, st):
    """One displayed waterfall line from n uint8 lines: a1 -> a2 -> a3 -> pixel (utils:879-893)."""
    frames = np.asarray(frames_u8, dtype=np.uint8).astype(np.float32)
    spectrum = time_bin_mean(frames) if frames.ndim == 2 and frames.shape[0] > 1 else frames.reshape(-1)
    colour = spectrum_db2col(spectrum, st)
    return spectrum, colour, pixel_row(colour)


def scroll(wf_data, row):
    """utils_supersdr.py:896-897 -- scroll the float64 image one line down, newest row on top."""
    wf_data[1:, :] = wf_data[0:-1, :]
    wf_data[0, :] = row
    return wf_data


def spectrum_trace(wf_data, spectrum_height, t_avg=15):
    """utils_supersdr.py:1678-1679 -- mean of the newest t_avg rows -> y pixel per bin."""
    v = np.nanmean(wf_data.T[:, :t_avg], axis=1)
    return np.array([spectrum_height - 1 - int(x / 255 * spectrum_height) for x in v], dtype=np.int64)


def cutesdr_palette():
    """display_stuff.create_cm('cutesdr'), utils_supersdr.py:1391-1412 -- 255 float RGB triples."""
    cm = []
    for i in range(255):
        if i < 43:
            col = (0, 0, 255 * (i) / 43)
        if (i >= 43) and (i < 87):
            col = (0, 255 * (i - 43) / 43, 255)
        if (i >= 87) and (i < 120):
            col = (0, 255, 255 - (255 * (i - 87) / 32))
        if (i >= 120) and (i < 154):
            col = ((255 * (i - 120) / 33), 255, 0)
        if (i >= 154) and (i < 217):
            col = (255, 255 - (255 * (i - 154) / 62), 0)
        if i >= 217:
            col = (255, 0, 128 * (i - 217) / 38)
        cm.append(col)
    return np.array(cm, dtype=np.float64)


# ---------------------------------------------------------------------------------------------
# audio
# ---------------------------------------------------------------------------------------------

def snd_ingest(msg):
    """utils_supersdr.py:1065-1072 -- returns (int16 samples, rssi dBm, adc_overflow)."""
    if msg[0:3] != b"SND":
        return None
    flags, seq = struct.unpack("<BI", bytes(msg[3:8]))
    s_meter, = struct.unpack(">H", bytes(msg[8:10]))
    rssi = 0.1 * s_meter - 127
    data = bytes(msg[10:])
    count = len(data) // 2
    samples = np.frombuffer(data, dtype=">h", count=count).astype(np.int16)
    return samples, rssi, bool(flags & 2), seq


def iq_ingest(msg):
    """kiwi/client.py:385-388,443-454 -- 'SND' frame in mod=iq: 10-byte prefix, <BBII GPS header,
    then big-endian int16 I,Q pairs -> complex64 (unscaled counts)."""
    flags, seq = struct.unpack("<BI", bytes(msg[3:8]))
    smeter, = struct.unpack(">H", bytes(msg[8:10]))
    data = bytes(msg[10:])
    gps = struct.unpack("<BBII", data[0:10])
    data = data[10:]
    count = len(data) // 2
    samples = np.frombuffer(data, dtype=">h", count=count).astype(np.float32)
    cs = np.empty(count // 2, dtype=np.complex64)
    cs.real = samples[0:count:2]
    cs.imag = samples[1:count:2]
    return cs, 0.1 * smeter - 127, gps, seq


def fir_design(fl, fs):
    """``filtering.__init__`` utils_supersdr.py:333-344 -- Blackman-windowed sinc, unity DC gain."""
    b = fl / fs
    N = int(np.ceil((4 / b)))
    if not N % 2:
        N += 1
    h = np.sinc(2. * fl / fs * (np.arange(N) - (N - 1) / 2.))
    h = h * np.blackman(N)
    return h / np.sum(h)


def lowpass(signal, h):
    """``filtering.lowpass`` utils_supersdr.py:346-348."""
    return np.convolve(signal, h, mode="valid")


class InterpState:
    """The attributes ``play_buffer`` reads on ``kiwi_sound`` (integer-ratio path)."""

    def __init__(self, kiwi_rate=12000, audio_rate=48000):
        self.ratio = audio_rate / kiwi_rate
        self.h = fir_design(kiwi_rate / 2, audio_rate)
        self.n_tap = len(self.h)
        self.old_buffer = np.zeros(self.n_tap - 1)
        g = np.gcd(kiwi_rate, audio_rate)
        self.n_low, self.n_high = int(kiwi_rate / g), int(audio_rate / g)


def play_buffer(popped_i16, st, volume=100, balance=0.0):
    """utils_supersdr.py:1121-1138.  Returns (float64 pre-cast mono buffer, int16[frames, 2])."""
    popped = np.asarray(popped_i16).flatten().astype(np.float64) * (volume / 100)
    n = len(popped)
    if st.ratio % 1:
        from scipy.signal import resample_poly
        buf = resample_poly(popped, st.n_high, st.n_low, padtype="line")[:-1]
    else:
        r = int(st.ratio)
        buf = np.zeros(r * n)
        buf[::r] = popped
        buf = np.concatenate([st.old_buffer, buf])
        st.old_buffer = buf[-(st.n_tap - 1):]
        buf = lowpass(buf, st.h) * r
    lv, rv = min(1 - balance, 1.0), min(1 + balance, 1.0)
    out = np.empty((len(buf), 2), dtype=np.int16)
    with np.errstate(invalid="ignore"):
        out[:, 0] = (buf * lv ** 2).astype(np.int16)
        out[:, 1] = (buf * rv ** 2).astype(np.int16)
    return buf, out


def mute_step(rssi, mute_counter, max_rssi_before_mute=-20, muting_delay=15):
    """utils_supersdr.py:1141-1147 -- returns (new counter, muted?)."""
    if rssi > max_rssi_before_mute:
        mute_counter = muting_delay
    elif mute_counter > 0:
        mute_counter -= 1
    return mute_counter, mute_counter > 0
