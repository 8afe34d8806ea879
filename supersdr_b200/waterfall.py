"""Waterfall side of the hot path.

``WaterfallBank`` is the batched engine (B independent channels per GPU) over the C ABI
(``ssdr_wf_*`` in include/ssdr_b200.h).  ``kiwi_waterfall`` keeps the attribute / method surface of
the reference class (utils_supersdr.py:592-898) so ``supersdr.py`` can use it unchanged; it owns a
one-channel bank and computes locally, from raw IQ frames, what the reference receives finished
from the KiwiSDR server (uint8 W/F lines, utils_supersdr.py:780-785) and then post-processes
(averaging :881-886, ``spectrum_db2col`` :787-813).
"""
import ctypes as C
from collections import deque

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


class kiwi_waterfall:
    """Drop-in for utils_supersdr.kiwi_waterfall (utils_supersdr.py:592-898).

    Same constructor signature and public attributes; ``iq_source`` replaces the W/F websocket: any
    object whose ``read_wf_frame()`` returns one frame of ``WF_BINS`` complex64 IQ samples (or
    ``None`` when the stream ended).  Everything ``supersdr.py`` reads (``wf_data``, ``wf_color``,
    ``spectrum``, ``wf_min_db`` ...) is produced by the CUDA path.
    """
    MAX_FREQ = 30000
    CENTER_FREQ = int(MAX_FREQ / 2)
    MAX_ZOOM = 14
    WF_BINS = 1024
    MAX_FPS = 23
    MIN_DYN_RANGE = 40.
    CLIP_LOWP, CLIP_HIGHP = 40., 100
    delta_low_db, delta_high_db = 0, 0
    low_clip_db, high_clip_db = -120, -60
    wf_min_db, wf_max_db = low_clip_db, low_clip_db + MIN_DYN_RANGE
    kiwi_wf_timestamp = None
    wf_buffer_len = 3

    def __init__(self, host_, port_, pass_, zoom_, freq_, eibi, disp, iq_source=None, wf_bins=None):
        self.eibi = eibi
        self.host, self.port, self.password = host_, port_, pass_
        self.zoom = zoom_
        self.freq = freq_ if freq_ else 14200
        self.averaging_n = 1
        self.wf_auto_scaling = True
        if wf_bins:
            self.WF_BINS = int(wf_bins)
        self.BINS2PIXEL_RATIO = disp.DISPLAY_WIDTH / self.WF_BINS
        self.old_averaging_n = self.averaging_n
        self.dynamic_range = self.MIN_DYN_RANGE
        self.wf_white_flag = False
        self.terminate = False
        self.run_index = 0
        self.tune = self.freq
        self.radio_mode = "USB"
        self.span_khz = self.zoom_to_span()
        self.start_f_khz = self.start_freq()
        self.end_f_khz = self.end_freq()
        self.div_list, self.subdiv_list = [], []
        self.min_bin_spacing = 100
        self.space_khz = 10
        self.counter, self.actual_freq = self.start_frequency_to_counter(self.start_f_khz)
        self.wf_color = None
        self.freq_offset = 0
        self.iq_source = iq_source
        if iq_source is None:
            raise Exception("no IQ source")          # reference raises on a failed connect, utils:667
        self.bins_per_khz = self.WF_BINS / self.span_khz
        self.wf_data = np.zeros((disp.WF_HEIGHT, self.WF_BINS))
        self.wf_data_tmp = deque([], self.wf_buffer_len)
        self.avg_spectrum_deque = deque([], self.averaging_n)
        self._bank = None
        self._bank_n = None

    # ---- frequency / zoom arithmetic: utils_supersdr.py:747-777 --------------------------------
    def zoom_to_span(self):
        assert self.zoom >= 0 and self.zoom <= self.MAX_ZOOM
        self.span_khz = self.MAX_FREQ / 2 ** self.zoom
        return self.span_khz

    def start_frequency_to_counter(self, start_frequency_):
        assert start_frequency_ >= 0 and start_frequency_ <= self.MAX_FREQ
        self.counter = round(start_frequency_ / self.MAX_FREQ * 2 ** self.MAX_ZOOM * self.WF_BINS)
        start_frequency_ = self.counter * self.MAX_FREQ / self.WF_BINS / 2 ** self.MAX_ZOOM
        return self.counter, start_frequency_

    def start_freq(self):
        self.start_f_khz = self.freq - self.span_khz / 2
        return self.start_f_khz

    def end_freq(self):
        self.end_f_khz = self.freq + self.span_khz / 2
        return self.end_f_khz

    def offset_to_bin(self, offset_khz_):
        return self.WF_BINS / self.span_khz * offset_khz_

    def bins_to_khz(self, bins_):
        return (1. / (self.WF_BINS / self.span_khz)) * bins_ + self.start_f_khz

    def deltabins_to_khz(self, bins_):
        return (1. / (self.WF_BINS / self.span_khz)) * bins_

    def gen_div(self):
        """Frequency-axis ticks as waterfall bin indices (behaviour of utils_supersdr.py:696-717):
        major ticks every ``space_khz`` (x10 until at least ``min_bin_spacing`` bins apart), minor ticks
        at a tenth of that; the spacing is escalated until one of the lists is non-empty."""
        self.space_khz = 10
        self.div_list, self.subdiv_list = [], []
        lo, hi = int(self.start_f_khz), int(self.end_f_khz)
        to_bin = lambda f: int(self.offset_to_bin(f - self.start_f_khz))
        while not self.div_list and not self.subdiv_list:
            minor = self.space_khz / 10
            if self.bins_per_khz * self.space_khz > self.min_bin_spacing:
                self.div_list = [to_bin(f) for f in range(lo, hi + 1) if not f % self.space_khz]
            if self.bins_per_khz * minor > self.min_bin_spacing / 10:
                self.subdiv_list = [to_bin(f) for f in range(lo, hi + 1) if not f % minor]
            self.space_khz *= 10

    def set_freq_zoom(self, freq_, zoom_):
        """utils_supersdr.py:815-845 (the SET zoom/start message becomes a source retune)."""
        self.freq, self.zoom = freq_, zoom_
        self.zoom_to_span(); self.start_freq(); self.end_freq()
        if zoom_ == 0:
            self.freq = self.CENTER_FREQ
            self.start_freq(); self.end_freq()
            self.span_khz = self.MAX_FREQ
        else:
            if self.start_f_khz < 0:
                self.freq = self.zoom_to_span() / 2
                self.start_freq(); self.end_freq(); self.zoom_to_span()
            elif self.end_f_khz > self.MAX_FREQ:
                self.freq = self.MAX_FREQ - self.zoom_to_span() / 2
                self.start_freq(); self.end_freq(); self.zoom_to_span()
        self.counter, actual_freq = self.start_frequency_to_counter(self.start_f_khz)
        if hasattr(self.iq_source, "set_zoom_start"):
            self.iq_source.set_zoom_start(self.zoom, self.counter)
        if self.eibi is not None and hasattr(self.eibi, "get_stations"):
            self.eibi.get_stations(self.start_f_khz, self.end_f_khz)
        self.bins_per_khz = self.WF_BINS / self.span_khz
        self.gen_div()
        return self.freq

    def change_passband(self, delta_low_, delta_high_):
        """utils_supersdr.py:859-873."""
        if self.radio_mode == "USB":
            lc_, hc_ = LOW_CUT_SSB + delta_low_, HIGH_CUT_SSB + delta_high_
        elif self.radio_mode == "LSB":
            lc_, hc_ = -HIGH_CUT_SSB - delta_high_, -LOW_CUT_SSB - delta_low_
        elif self.radio_mode == "AM":
            lc_, hc_ = -HIGHLOW_CUT_AM - delta_low_, HIGHLOW_CUT_AM + delta_high_
        elif self.radio_mode == "CW":
            lc_, hc_ = LOW_CUT_CW + delta_low_, HIGH_CUT_CW + delta_high_
        self.lc, self.hc = lc_, hc_
        return lc_, hc_

    def keepalive(self):
        if hasattr(self.iq_source, "keepalive"):
            self.iq_source.keepalive()

    def close_connection(self):
        if hasattr(self.iq_source, "close"):
            self.iq_source.close()

    def set_white_flag(self):
        self.wf_color = np.ones_like(self.wf_color) * 255
        self.wf_data[0, :] = self.wf_color

    # ---- the hot path ----------------------------------------------------------------------------
    def _get_bank(self, n):
        if self._bank is None or self._bank_n != n:
            if self._bank is not None:
                self._bank.close()
            self._bank = WaterfallBank(self.WF_BINS, 1, n)
            self._bank_n = n
        return self._bank

    def receive_spectrum(self):
        """utils_supersdr.py:780-785: one line.  Here: one IQ frame -> FFT -> Kiwi byte line (float32)."""
        frame = self.iq_source.read_wf_frame()
        if frame is None:
            self.terminate = True
            return None
        self._frames.append(np.asarray(frame, dtype=np.complex64).reshape(self.WF_BINS))
        self.keepalive()
        return frame

    def spectrum_db2col(self):
        """utils_supersdr.py:787-813 on the GPU: consumes the frames gathered by receive_spectrum."""
        n = len(self._frames)
        bank = self._get_bank(n)
        bank.set_display(0, 1, zoom=self.zoom, auto_scale=self.wf_auto_scaling, delta_low_db=self.delta_low_db,
                         delta_high_db=self.delta_high_db, low_clip_db=float(self.low_clip_db),
                         dynamic_range=float(self.dynamic_range))
        res = bank.process(np.stack(self._frames)[None, :, :])
        sc = res["scalars"][0]
        self.spectrum = res["spectrum"][0]
        self.wf_color = res["colour"][0]
        self.wf_pixels = res["pixels"][0]
        if self.wf_auto_scaling:
            self.low_clip_db = sc["low_clip_db"]
            self.high_clip_db = sc["high_clip_db"]
            self.dynamic_range = sc["dynamic_range"]
        self.wf_min_db = sc["wf_min_db"]
        self.wf_max_db = sc["wf_max_db"]

    def run_once(self):
        """One iteration of ``run`` (utils_supersdr.py:880-897)."""
        self._frames = []
        n = self.averaging_n if self.averaging_n > 1 else 1
        for _ in range(n):
            if self.receive_spectrum() is None:
                return False
        self.run_index += 1
        self.spectrum_db2col()
        self.wf_data_tmp.appendleft(self.wf_color)
        if len(self.wf_data_tmp) > 0 and self.run_index > self.wf_buffer_len:
            self.wf_data[1:, :] = self.wf_data[0:-1, :]
            self.wf_data[0, :] = self.wf_data_tmp.pop()
        return True

    def run(self):
        while not self.terminate:
            if not self.run_once():
                break
        return
