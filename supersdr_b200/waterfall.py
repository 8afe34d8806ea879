"""Waterfall side of the hot path.

``WaterfallBank`` is the batched engine (B independent channels per GPU) over the C ABI
(``ssdr_wf_*`` in include/ssdr_b200.h): from raw IQ frames it computes what the reference receives
finished from the KiwiSDR server (uint8 W/F lines, utils_supersdr.py:780-785) and then post-processes
(averaging :881-886, ``spectrum_db2col`` :787-813); ``colorrow`` is the entry for finished lines.
``WaterfallImage`` is the scrolling image / palette / trace on the GPU.  The drop-in for the reference
class ``kiwi_waterfall`` is made by delegation in ``supersdr_b200.dropin``.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib, check, ptr

# mode pass-bands, utils_supersdr.py:42-50
LOW_CUT_SSB, HIGH_CUT_SSB = 30, 3000
CW_PITCH = 0.6
LOW_CUT_CW, HIGH_CUT_CW = int(CW_PITCH * 1000 - 200), int(CW_PITCH * 1000 + 200)
HIGHLOW_CUT_AM = 6000


def _percentile_index_formula(n, q_percent):
    """numpy >= 2.0's expression: ``(n - 1) * (q / float32(100))`` evaluated in float32."""
    q = np.float32(q_percent) / np.float32(100)
    vi = np.float32(n - 1) * q
    lo = int(np.floor(vi))
    gamma = np.float32(vi - np.float32(lo))
    if lo >= n - 1:
        lo, gamma = n - 1, np.float32(0)
    return lo, float(gamma)


_PCT_CACHE = {}


def percentile_index(n, q_percent=40.0):
    """(lo, gamma) of ``np.percentile(x_float32[n], q)`` ('linear' method) AS THE INSTALLED numpy EVALUATES IT, so that
    the kernel's ``s[lo] + (s[lo+1] - s[lo]) * gamma`` reproduces the reference's ``np.percentile(wf_db, 40.)``
    (utils_supersdr.py:794) bit for bit whatever the numpy version: the pair is measured by probing ``np.percentile``
    with step vectors (0 up to index j, 1 above: the result is 0, gamma or 1 according to where j lies) around the
    analytic value of numpy >= 2.0.  A gamma that is not a float32 number (numpy 1.x interpolates in float64) cannot
    be reproduced by the float32 kernel and is refused."""
    key = (int(n), float(q_percent))
    if key in _PCT_CACHE:
        return _PCT_CACHE[key]
    lo0, g0 = _percentile_index_formula(n, q_percent)
    found = None
    if n >= 2:
        idx = np.arange(n)
        probe = lambda j: float(np.percentile((idx > j).astype(np.float32), q_percent))
        for j in sorted(range(max(lo0 - 2, 0), min(lo0 + 3, n - 1)), key=lambda j: abs(j - lo0)):
            r = probe(j)
            if 0.0 < r < 1.0:                        # j is the lower index, r the weight
                found = (j, r)
                break
            if r == 0.0 and (j == 0 or probe(j - 1) == 1.0):   # weight 0: pure selection of s[j]
                found = (j, 0.0)
                break
    if found is None:
        found = (lo0, g0)
    if float(np.float32(found[1])) != found[1]:
        raise RuntimeError("numpy %s interpolates percentiles in float64 (weight %r): the float32 colour row of "
                           "utils_supersdr.py:794 cannot be matched bit for bit; use numpy >= 2.0" % (np.__version__, found[1]))
    _PCT_CACHE[key] = found
    return found


class WaterfallBank:
    """B channels x n_avg frames x nfft IQ samples  ->  one colour row per channel."""

    CLIP_LOWP = 40.0          # kiwi_waterfall.CLIP_LOWP, utils_supersdr.py:599

    def __init__(self, nfft=1024, batch=1, n_avg=1, window=True, cal_db=_lib.WF_CAL_DB, device=None):
        _lib.init(device)
        self.nfft, self.batch, self.n_avg = int(nfft), int(batch), int(n_avg)
        self.window, self.cal_db = bool(window), float(cal_db)
        lo, gamma = percentile_index(self.nfft, self.CLIP_LOWP)
        h = C.c_void_p()
        check(lib.ssdr_wf_create(C.byref(h), self.nfft, self.batch, self.n_avg, int(self.window),
                                 self.cal_db, lo, gamma))
        self._h = h

    # -- parameters ---------------------------------------------------------------------------
    def set_display(self, first=0, count=None, zoom=0, auto_scale=True, delta_low_db=0, delta_high_db=0,
                    low_clip_db=-120.0, dynamic_range=40.0):
        """Per-channel attributes that ``spectrum_db2col`` reads (utils_supersdr.py:592-620)."""
        count = self.batch - first if count is None else count
        arr = (_lib.WfDisplay * count)()
        for i in range(count):
            arr[i] = _lib.WfDisplay(int(zoom), int(bool(auto_scale)), int(delta_low_db), int(delta_high_db),
                                    float(low_clip_db), float(dynamic_range))
        check(lib.ssdr_wf_set_display(self._h, first, count, arr))

    def set_remote_input(self, remote=True):
        """Device inputs of ``process_dev`` / ``time_dev`` live in a peer GPU's memory (sharding.open_peer_buffer)."""
        check(lib.ssdr_wf_set_remote_input(self._h, int(bool(remote))))

    def tables(self):
        """(twiddles complex64[nfft], thresholds float32[256], radix plan) the kernels use."""
        tw = np.empty(2 * self.nfft, np.float32)
        th = np.empty(256, np.float32)
        r = (C.c_int * 8)()
        n = check(lib.ssdr_wf_get_tables(self._h, ptr(tw), ptr(th), r))
        return tw.view(np.complex64), th, [r[i] for i in range(n)]

    def window_table(self):
        """First half of the periodic Hann window the kernel uses, float32[nfft / 2]."""
        w = np.empty(self.nfft // 2, np.float32)
        check(lib.ssdr_wf_get_window(self._h, ptr(w)))
        return w

    # -- host-buffer API ----------------------------------------------------------------------
    def _outputs(self, want_colour, want_spectrum, out):
        B, N = self.batch, self.nfft
        out = {} if out is None else out
        if "pixels" not in out:
            out["pixels"] = np.empty((B, N), np.uint8)
        if want_colour and "colour" not in out:
            out["colour"] = np.empty((B, N), np.float32)
        if want_spectrum and "spectrum" not in out:
            out["spectrum"] = np.empty((B, N), np.float32)
        if "scalars" not in out:
            out["scalars"] = np.empty(B, _lib.SCALARS_DTYPE)
        return out

    def process(self, iq, want_colour=True, want_spectrum=True, out=None):
        """iq: complex64[B, n_avg, nfft] (int16-count units) or uint8[B, n_avg, nfft, 4] big-endian
        int16 I,Q wire bytes.  Returns dict(pixels, colour, spectrum, scalars)."""
        iq = np.asarray(iq)
        if iq.dtype == np.complex64:
            fmt = _lib.SSDR_IQ_CF32
            expect = (self.batch, self.n_avg, self.nfft)
        elif iq.dtype == np.uint8:
            fmt = _lib.SSDR_IQ_S16BE
            expect = (self.batch, self.n_avg, self.nfft, 4)
        else:
            raise TypeError("iq must be complex64 or uint8 wire bytes, got %s" % iq.dtype)
        if iq.shape != expect:
            raise ValueError("iq shape %s, expected %s" % (iq.shape, expect))
        iq = np.ascontiguousarray(iq)
        out = self._outputs(want_colour, want_spectrum, out)
        check(lib.ssdr_wf_process(self._h, ptr(iq), fmt, ptr(out["pixels"]), ptr(out.get("colour")),
                                  ptr(out.get("spectrum")), ptr(out["scalars"])))
        return out

    def colorrow(self, lines_u8, want_colour=True, want_spectrum=True, out=None):
        """Tier-P entry: finished Kiwi W/F lines uint8[B, n_avg, nfft] (utils_supersdr.py:783-784)
        -> mean + spectrum_db2col + pixels."""
        lines = np.ascontiguousarray(lines_u8, dtype=np.uint8)
        if lines.shape != (self.batch, self.n_avg, self.nfft):
            raise ValueError("lines shape %s, expected %s" % (lines.shape, (self.batch, self.n_avg, self.nfft)))
        out = self._outputs(want_colour, want_spectrum, out)
        check(lib.ssdr_wf_colorrow_u8(self._h, ptr(lines), ptr(out["pixels"]), ptr(out.get("colour")),
                                      ptr(out.get("spectrum")), ptr(out["scalars"])))
        return out

    # -- device-resident API (bench, multi-GPU) -----------------------------------------------
    def process_dev(self, iq_dev, fmt, pixels_dev, colour_dev=None, spectrum_dev=None, scalars_dev=None):
        check(lib.ssdr_wf_process_dev(self._h, iq_dev, fmt, pixels_dev, colour_dev, spectrum_dev, scalars_dev))

    def time_dev(self, iq_dev, fmt, pixels_dev, iters=1):
        ms = C.c_float()
        check(lib.ssdr_wf_time_dev(self._h, iq_dev, fmt, pixels_dev, int(iters), C.byref(ms)))
        return ms.value

    def sync(self):
        check(lib.ssdr_wf_sync(self._h))

    def close(self):
        if getattr(self, "_h", None):
            lib.ssdr_wf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def create_cm(which="cutesdr"):
    """display_stuff.create_cm (utils_supersdr.py:1391-1412): the 255-entry CuteSDR colour map, as float tuples."""
    colormap = []
    if which == "cutesdr":
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
            colormap.append(col)
    return colormap


def palette_u8(colormap):
    """uint8[256][3] look-up table from a ``create_cm`` colour map: components truncated to integers (what an 8-bit
    palette stores; the conversion happens inside pygame -- parity unpinned), entry 255 = white (the value
    ``set_white_flag`` writes, utils_supersdr.py:875-877)."""
    pal = np.full((256, 3), 255, np.uint8)
    cm = np.asarray(colormap, dtype=np.float64)
    pal[:len(cm)] = np.clip(np.trunc(cm), 0, 255).astype(np.uint8)
    return pal


class WaterfallImage:
    """The scrolling waterfall image and the spectrum trace on the GPU: ``push`` takes one
    ``wf_color`` row per channel and reproduces kiwi_waterfall.run's 3-deep delay deque and one-line scroll
    (utils_supersdr.py:893-897) on a ring buffer; ``image`` returns RGB through the palette and/or ``wf_data``;
    ``trace`` is display_stuff.plot_spectrum's ``nanmean`` of the newest 15 lines and its y pixel (:1678-1679)."""

    def __init__(self, batch, height, width, colormap="cutesdr", device=None):
        _lib.init(device)
        self.batch, self.height, self.width = int(batch), int(height), int(width)
        self.palette = palette_u8(create_cm(colormap) if isinstance(colormap, str) else colormap)
        h = C.c_void_p()
        check(lib.ssdr_wf_image_create(C.byref(h), self.batch, self.height, self.width, ptr(np.ascontiguousarray(self.palette))))
        self._h = h

    def push(self, colour_rows):
        rows = np.ascontiguousarray(colour_rows, dtype=np.float32)
        if rows.shape != (self.batch, self.width):
            raise ValueError("colour rows must be float32[%d, %d]" % (self.batch, self.width))
        check(lib.ssdr_wf_image_push(self._h, ptr(rows)))

    def push_dev(self, colour_dev_ptr):
        check(lib.ssdr_wf_image_push_dev(self._h, colour_dev_ptr))

    def set_white_flag(self):
        check(lib.ssdr_wf_image_white(self._h))

    def image(self, want_rgb=True, want_data=False):
        rgb = np.empty((self.batch, self.height, self.width, 3), np.uint8) if want_rgb else None
        data = np.empty((self.batch, self.height, self.width), np.float64) if want_data else None
        check(lib.ssdr_wf_image_get(self._h, ptr(rgb), ptr(data)))
        return rgb, data

    def trace(self, t_avg=15, spectrum_height=200):
        v = np.empty((self.batch, self.width), np.float64)
        y = np.empty((self.batch, self.width), np.int32)
        check(lib.ssdr_wf_image_trace(self._h, int(t_avg), int(spectrum_height), ptr(v), ptr(y)))
        return v, y

    def close(self):
        if getattr(self, "_h", None):
            lib.ssdr_wf_image_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
