"""Audio side of the hot path.

``DemodBank`` (K4) and ``InterpBank`` (K5) are the batched engines over the C ABI
(``ssdr_demod_*`` / ``ssdr_interp_*`` in include/ssdr_b200.h).  ``filtering`` keeps the surface of the
reference class (utils_supersdr.py:333-348).  The demodulator that the reference leaves to the remote
KiwiSDR (it only sends ``SET mod=... low_cut=... high_cut=...`` and ``SET agc=...``,
utils_supersdr.py:1022-1029) runs here on the GPU from raw IQ, followed by the reference's own x4
interpolation filter.  The drop-in for the reference class ``kiwi_sound`` is made by delegation in
``supersdr_b200.dropin``.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, ptr
from .waterfall import (LOW_CUT_SSB, HIGH_CUT_SSB, LOW_CUT_CW, HIGH_CUT_CW, HIGHLOW_CUT_AM)

MODE_IDS = {"am": _lib.MODE_AM, "usb": _lib.MODE_USB, "lsb": _lib.MODE_LSB, "cw": _lib.MODE_CW,
            "nbfm": _lib.MODE_NBFM}


def default_passband(mode):
    """SuperSDR pass-bands (utils_supersdr.py:42-50,859-873); NBFM from kiwi/client.py:237-239."""
    return {"usb": (LOW_CUT_SSB, HIGH_CUT_SSB), "lsb": (-HIGH_CUT_SSB, -LOW_CUT_SSB),
            "am": (-HIGHLOW_CUT_AM, HIGHLOW_CUT_AM), "cw": (LOW_CUT_CW, HIGH_CUT_CW),
            "nbfm": (-6000, 6000)}[mode.lower()]


def design_lowpass(fl, fs, n_taps=None):
    """Blackman-windowed sinc with unity DC gain -- the recipe of ``filtering.__init__``
    (utils_supersdr.py:333-344).  ``n_taps=None`` uses the reference's length rule."""
    if n_taps is None:
        n_taps = int(np.ceil(4 / (fl / fs)))
        if not n_taps % 2:
            n_taps += 1
    k = np.arange(n_taps) - (n_taps - 1) / 2.
    h = np.sinc(2. * fl / fs * k) * np.blackman(n_taps)
    return h / np.sum(h)


def demod_params(mode="usb", lc=None, hc=None, f_off=0.0, on=True, hang=False, thresh=-80, slope=0,
                 decay=4000, gain=50):
    """Build one ``ssdr_demod_params_t``.  AGC names / defaults: utils_supersdr.py:936-945."""
    mode = mode.lower()
    dlc, dhc = default_passband(mode)
    lc = dlc if lc is None else lc
    hc = dhc if hc is None else hc
    p = _lib.DemodParams()
    p.mode = MODE_IDS[mode]
    p.low_cut_hz, p.high_cut_hz, p.freq_offset_hz = float(lc), float(hc), float(f_off)
    p.agc_on, p.agc_hang = int(bool(on)), int(bool(hang))
    p.agc_thresh_dbm, p.agc_slope_db, p.agc_decay_ms, p.agc_man_gain_db = float(thresh), float(slope), float(decay), float(gain)
    taps = design_lowpass((hc - lc) / 2.0, _lib.KIWI_RATE, _lib.FIR_TAPS)
    taps[0] = taps[-1] = 0.0      # Blackman end points are exactly zero, not +-1e-17
    taps = taps.astype(np.float32)
    for i in range(_lib.FIR_TAPS):
        p.taps[i] = float(taps[i])
    return p


def demod_plan(params, n_sm=148):
    """Work plan of the tcgen05 FIR engine for a list of ``demod_params`` (host only, no device needed): returns
    ``(quad_ch int32[n_quads, 4], quad_fid int32[n_quads], tiles_per_round, fill)`` -- see ``ssdr_demod_plan``."""
    B = len(params)
    taps = np.ascontiguousarray([[p.taps[i] for i in range(_lib.FIR_TAPS)] for p in params], np.float32)
    work = np.ascontiguousarray([p.mode * 4 + p.agc_on * 2 + p.agc_hang for p in params], np.int32)
    n, tiles, fill = C.c_int(), C.c_int(), C.c_float()
    check(lib.ssdr_demod_plan(ptr(taps), ptr(work), B, int(n_sm), None, None, 0, C.byref(n), C.byref(tiles), C.byref(fill)))
    qc, qf = np.empty((n.value, 4), np.int32), np.empty(n.value, np.int32)
    check(lib.ssdr_demod_plan(ptr(taps), ptr(work), B, int(n_sm), ptr(qc), ptr(qf), n.value, C.byref(n), C.byref(tiles), C.byref(fill)))
    return qc, qf, tiles.value, fill.value


class DemodBank:
    """B channels of IQ @12 kHz -> PCM @12 kHz (float32 + int16) and per-frame RSSI."""

    ENGINES = {"ffma": _lib.SSDR_DEMOD_ENGINE_FFMA, "tcgen05": _lib.SSDR_DEMOD_ENGINE_TCGEN05, "auto": _lib.SSDR_DEMOD_ENGINE_AUTO}

    def __init__(self, batch=1, max_samples=_lib.FRAME * 64, device=None, engine=None):
        _lib.init(device)
        self.batch, self.max_samples = int(batch), int(max_samples)
        h = C.c_void_p()
        check(lib.ssdr_demod_create(C.byref(h), self.batch, self.max_samples))
        self._h = h
        if engine is not None:
            self.set_engine(engine)
        self.set_params(0, [demod_params()] * self.batch)

    def set_params(self, first, params):
        arr = (_lib.DemodParams * len(params))(*params)
        check(lib.ssdr_demod_set(self._h, int(first), len(params), arr))

    def set_engine(self, engine):
        """FIR engine of the fused kernel: "ffma" (fp32 pipe), "tcgen05" (tensor cores) or "auto" (default: tensor cores
        when channels share filters); switchable between calls."""
        check(lib.ssdr_demod_set_engine(self._h, self.ENGINES[engine] if isinstance(engine, str) else int(engine)))

    def set_all(self, **kw):
        self.set_params(0, [demod_params(**kw)] * self.batch)

    def reset(self):
        check(lib.ssdr_demod_reset(self._h))

    def process(self, iq, want_f32=True, want_i16=True, want_rssi=True, out=None):
        """iq: complex64[B, n] or uint8[B, n, 4] wire bytes; n a multiple of 512.  ``out`` may hold caller-owned
        (ideally pinned) result arrays under the keys ``pcm_f32`` / ``pcm_i16`` / ``rssi``."""
        iq = np.asarray(iq)
        if iq.dtype == np.complex64:
            fmt, n = _lib.SSDR_IQ_CF32, iq.shape[-1]
            ok = iq.shape == (self.batch, n)
        elif iq.dtype == np.uint8:
            fmt, n = _lib.SSDR_IQ_S16BE, iq.shape[1]
            ok = iq.shape == (self.batch, n, 4)
        else:
            raise TypeError("iq must be complex64 or uint8 wire bytes, got %s" % iq.dtype)
        if not ok:
            raise ValueError("iq shape %s does not match batch %d" % (iq.shape, self.batch))
        iq = np.ascontiguousarray(iq)
        out = out or {}

        def buf(key, want, shape, dtype):
            if not want:
                return None
            a = out.get(key)
            if a is None:
                return np.empty(shape, dtype)
            if a.shape != shape or a.dtype != dtype or not a.flags.c_contiguous:
                raise ValueError("out[%r] must be a C-contiguous %s array of shape %s" % (key, np.dtype(dtype).name, shape))
            return a
        f32 = buf("pcm_f32", want_f32, (self.batch, n), np.float32)
        i16 = buf("pcm_i16", want_i16, (self.batch, n), np.int16)
        rssi = buf("rssi", want_rssi, (self.batch, n // _lib.FRAME), np.float32)
        check(lib.ssdr_demod_process(self._h, ptr(iq), fmt, int(n), ptr(f32), ptr(i16), ptr(rssi)))
        return dict(pcm_f32=f32, pcm_i16=i16, rssi=rssi)

    def process_dev(self, iq_dev, fmt, n, f32_dev=None, i16_dev=None, rssi_dev=None):
        check(lib.ssdr_demod_process_dev(self._h, iq_dev, fmt, int(n), f32_dev, i16_dev, rssi_dev))

    def time_dev(self, iq_dev, fmt, n, f32_dev=None, i16_dev=None, iters=1):
        ms = C.c_float()
        check(lib.ssdr_demod_time_dev(self._h, iq_dev, fmt, int(n), f32_dev, i16_dev, int(iters), C.byref(ms)))
        return ms.value

    def sync(self):
        check(lib.ssdr_demod_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib.ssdr_demod_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class InterpBank:
    """B channels of int16 PCM -> zero-stuff x ratio + FIR -> stereo int16 (play_buffer's
    integer-ratio path, utils_supersdr.py:1121-1138)."""

    def __init__(self, batch=1, ratio=4, taps=None, max_samples=_lib.FRAME * 64, device=None):
        _lib.init(device)
        self.batch, self.ratio, self.max_samples = int(batch), int(ratio), int(max_samples)
        if taps is None:
            taps = design_lowpass(_lib.KIWI_RATE / 2, _lib.KIWI_RATE * self.ratio)
        self.taps = np.ascontiguousarray(taps, dtype=np.float64)
        h = C.c_void_p()
        check(lib.ssdr_interp_create(C.byref(h), self.batch, self.ratio, ptr(self.taps), self.taps.size, self.max_samples))
        self._h = h

    def reset(self):
        check(lib.ssdr_interp_reset(self._h))

    def process(self, pcm_i16, volume=100, balance=0.0, want_mono=False):
        pcm = np.ascontiguousarray(pcm_i16, dtype=np.int16)
        if pcm.ndim != 2 or pcm.shape[0] != self.batch:
            raise ValueError("pcm shape %s does not match batch %d" % (pcm.shape, self.batch))
        n = pcm.shape[1]
        vol = np.ascontiguousarray(np.broadcast_to(np.asarray(volume, np.float32), (self.batch,)))
        bal = np.ascontiguousarray(np.broadcast_to(np.asarray(balance, np.float32), (self.batch,)))
        out = np.empty((self.batch, n * self.ratio, 2), np.int16)
        mono = np.empty((self.batch, n * self.ratio), np.float64) if want_mono else None
        check(lib.ssdr_interp_process(self._h, ptr(pcm), int(n), ptr(vol), ptr(bal), ptr(out), ptr(mono)))
        return (out, mono) if want_mono else out

    def close(self):
        if getattr(self, "_h", None):
            lib.ssdr_interp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ResampleLine:
    """``scipy.signal.resample_poly(x, up, down, padtype="line")[:-1]`` on the GPU -- kiwi_sound.play_buffer's path for
    non-integer sample ratios (utils_supersdr.py:1125-1126).  The polyphase filter is designed on the host exactly as
    resample_poly designs it (scipy.signal.firwin, Kaiser beta 5, 10 * max(up, down) taps per side, gain up, zero-padded
    in front so that output samples sit at the filter centre); the block is stateless, like the reference."""

    def __init__(self, up, down, device=None):
        from math import gcd
        from scipy.signal import firwin          # host-side filter design only, as the reference's resample_poly call
        _lib.init(device)
        g = gcd(int(up), int(down))
        self.up, self.down = int(up) // g, int(down) // g
        max_rate = max(self.up, self.down)
        self.half_len = 10 * max_rate
        h = firwin(2 * self.half_len + 1, 1.0 / max_rate, window=("kaiser", 5.0)).astype(np.float64)
        h *= self.up
        self.n_pre_pad = self.down - self.half_len % self.down
        self.n_pre_remove = (self.half_len + self.n_pre_pad) // self.down
        self._h_design = h

    def _plan(self, n_in):
        n_out = n_in * self.up
        n_out = n_out // self.down + bool(n_out % self.down)
        olen = lambda len_h: (((n_in - 1) * self.up + len_h) - 1) // self.down + 1
        n_post_pad = 0
        while olen(len(self._h_design) + self.n_pre_pad + n_post_pad) < n_out + self.n_pre_remove:
            n_post_pad += 1
        h = np.concatenate((np.zeros(self.n_pre_pad), self._h_design, np.zeros(n_post_pad)))
        return np.ascontiguousarray(h), n_out

    def process(self, pcm_i16, volume=100, balance=0.0, want_mono=False, drop_last=True):
        """pcm_i16: int16[B, n] -> stereo int16[B, n_keep, 2] (n_keep = resample_poly's length, minus one when
        ``drop_last`` as the reference does) and optionally the float64 mono buffer."""
        pcm = np.ascontiguousarray(pcm_i16, dtype=np.int16)
        if pcm.ndim != 2:
            raise ValueError("pcm must be int16[B, n]")
        B, n = pcm.shape
        h, n_out = self._plan(n)
        n_keep = n_out - 1 if drop_last else n_out
        vol = np.ascontiguousarray(np.broadcast_to(np.asarray(volume, np.float32), (B,)))
        bal = np.ascontiguousarray(np.broadcast_to(np.asarray(balance, np.float32), (B,)))
        out = np.empty((B, n_keep, 2), np.int16)
        mono = np.empty((B, n_keep), np.float64) if want_mono else None
        check(lib.ssdr_resample_line(ptr(pcm), B, n, ptr(vol), ptr(bal), ptr(h), int(h.size), self.up, self.down,
                                     int(self.n_pre_remove), int(n_keep), ptr(out), ptr(mono)))
        return (out, mono) if want_mono else out


class ImaAdpcmDecoder:
    """Batched drop-in for kiwi/client.py:58-87: ``decode(data)`` turns 4-bit IMA-ADPCM codes into int16 samples and
    carries (index, prev) from call to call.  ``batch`` independent streams decode in one launch."""

    def __init__(self, batch=1, device=None):
        _lib.init(device)
        self.batch = int(batch)
        self.state = np.zeros((self.batch, 2), np.int32)          # (index, prev) per stream

    @property
    def index(self):
        return int(self.state[0, 0])

    @property
    def prev(self):
        return int(self.state[0, 1])

    def decode(self, data):
        """bytes (one stream) or uint8[batch, n_bytes] -> int16 samples, two per byte (low nibble first)."""
        single = isinstance(data, (bytes, bytearray, memoryview))
        arr = np.frombuffer(bytes(data), np.uint8)[None] if single else np.ascontiguousarray(data, dtype=np.uint8)
        if arr.ndim != 2 or arr.shape[0] != self.batch:
            raise ValueError("data must be bytes or uint8[%d, n_bytes]" % self.batch)
        arr = np.ascontiguousarray(arr)
        out = np.empty((self.batch, 2 * arr.shape[1]), np.int16)
        check(lib.ssdr_adpcm_decode(ptr(arr), self.batch, int(arr.shape[1]), ptr(self.state), ptr(out)))
        return out[0] if single else out


class filtering:
    """Drop-in for utils_supersdr.filtering (utils_supersdr.py:333-348): same design, ``lowpass``
    evaluated on the GPU (a one-channel, ratio-1 interpolator bank = plain 'valid' FIR)."""

    def __init__(self, fl, fs):
        self.h = design_lowpass(fl, fs)
        self.n_tap = len(self.h)

    def lowpass(self, signal):
        """``np.convolve(signal, self.h, mode="valid")`` evaluated on the GPU in float64."""
        _lib.init()
        x = np.ascontiguousarray(signal, dtype=np.float64)
        out = np.empty(max(x.size - self.n_tap + 1, 0), np.float64)
        check(lib.ssdr_fir_valid_f64(ptr(x), x.size, ptr(np.ascontiguousarray(self.h)), self.n_tap, ptr(out)))
        return out


def unpack_iq(wire_bytes):
    """kiwi/client.py:449-453: big-endian int16 I,Q pairs -> complex64 (unscaled), on the GPU."""
    _lib.init()
    raw = np.ascontiguousarray(np.frombuffer(bytes(wire_bytes), dtype=np.uint8))
    n = raw.size // 4
    out = np.empty(n, np.complex64)
    check(lib.ssdr_unpack_iq_s16be(ptr(raw), ptr(out), n))
    return out
