"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper around oracle/_build/libssdr_oracle.so
(the scalar C restatement in oracle/c/ssdr_oracle.c).  Built by ``make -C oracle`` /
``__graft_entry__.build()``.  Never imported by ``supersdr_b200``."""
import ctypes as C
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import tier_p

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libssdr_oracle.so")
_lib = None


class ColourT(C.Structure):
    _fields_ = [("zoom", C.c_int), ("auto_scale", C.c_int), ("delta_low_db", C.c_int),
                ("delta_high_db", C.c_int), ("p_lo", C.c_int), ("p_gamma", C.c_float),
                ("low_clip_db", C.c_float), ("dynamic_range", C.c_float),
                ("high_clip_db", C.c_float), ("wf_min_db", C.c_float), ("wf_max_db", C.c_float)]


def build(force=False):
    if force or not os.path.isfile(_SO) or \
            os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "c", "ssdr_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.isfile(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.so_fft_plan.argtypes = [C.c_int, C.POINTER(C.c_int)]
        _lib.so_wf_frame_bytes.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        _lib.so_wf_rows.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.POINTER(ColourT), C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]
        _lib.so_twiddle_table.argtypes = [C.c_int, C.c_void_p]
        _lib.so_thresholds.argtypes = [C.c_int, C.c_double, C.c_void_p]
        _lib.so_window_table.argtypes = [C.c_int, C.c_void_p]
        _lib.so_markstein_mismatches.argtypes = [C.c_void_p, C.c_int, C.c_float]
        _lib.so_play_buffer.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_double, C.c_void_p,
                                        C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.so_colour_row.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ColourT),
                                       C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def fft_plan(N):
    r = (C.c_int * 8)()
    n = lib().so_fft_plan(N, r)
    if n < 0:
        raise ValueError("unsupported FFT size %d" % N)
    return [r[i] for i in range(n)]


def twiddle_table(N):
    t = np.empty(2 * N, np.float32)
    lib().so_twiddle_table(N, _p(t))
    return t.view(np.complex64)


def window_table(N):
    w = np.empty(N // 2, np.float32)
    lib().so_window_table(N, _p(w))
    return w


def markstein_mismatches(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return lib().so_markstein_mismatches(_p(a), a.size, float(np.float32(b)))


def thresholds(N, cal_db):
    t = np.empty(256, np.float32)
    lib().so_thresholds(N, cal_db, _p(t))
    return t


def wf_frame_bytes(iq, window=True, cal_db=-10.0, want_spectrum=False):
    """One frame complex64[N] -> (uint8[N] Kiwi bytes fftshifted, optional complex64 FFT)."""
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    N = iq.size
    out = np.empty(N, np.uint8)
    spec = np.empty(N, np.complex64) if want_spectrum else None
    if lib().so_wf_frame_bytes(_p(iq), N, int(window), cal_db, _p(out), _p(spec)):
        raise ValueError("unsupported FFT size %d" % N)
    return (out, spec) if want_spectrum else out


def colour_struct(W, zoom=0, auto_scale=True, delta_low_db=0, delta_high_db=0,
                  low_clip_db=-120.0, dynamic_range=40.0):
    lo, gamma = tier_p.percentile_virtual_index(W, tier_p.ColourState.CLIP_LOWP)
    return ColourT(zoom, int(auto_scale), delta_low_db, delta_high_db, lo, float(gamma),
                   low_clip_db, dynamic_range, 0.0, 0.0, 0.0)


def wf_rows(iq, window=True, cal_db=-10.0, threads=1, **colour_kw):
    """iq complex64[B, n, N] -> dict(pixels u8[B,N], colour f32, spectrum f32, sums u16, scalars f32[B,5])."""
    iq = np.ascontiguousarray(iq, dtype=np.complex64)
    B, n, N = iq.shape
    st = colour_struct(N, **colour_kw)
    px = np.empty((B, N), np.uint8)
    col = np.empty((B, N), np.float32)
    spec = np.empty((B, N), np.float32)
    sums = np.empty((B, N), np.uint16)
    sc = np.empty((B, 5), np.float32)

    def work(lohi):
        lo, hi = lohi
        if hi > lo:
            rc = lib().so_wf_rows(_p(iq[lo:hi]), hi - lo, n, N, int(window), cal_db, C.byref(st),
                                  _p(px[lo:hi]), _p(col[lo:hi]), _p(spec[lo:hi]), _p(sums[lo:hi]),
                                  _p(sc[lo:hi]))
            if rc:
                raise ValueError("unsupported FFT size %d" % N)
    threads = max(1, min(threads, B))
    edges = np.linspace(0, B, threads + 1).astype(int)
    if threads == 1:
        work((0, B))
    else:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(work, zip(edges[:-1], edges[1:])))
    return dict(pixels=px, colour=col, spectrum=spec, sums=sums, scalars=sc)


def colour_row(sums, n, **colour_kw):
    sums = np.ascontiguousarray(sums, dtype=np.uint16)
    W = sums.size
    st = colour_struct(W, **colour_kw)
    spec = np.empty(W, np.float32); col = np.empty(W, np.float32); px = np.empty(W, np.uint8)
    lib().so_colour_row(_p(sums), W, n, C.byref(st), _p(spec), _p(col), _p(px))
    return spec, col, px, np.array([st.low_clip_db, st.high_clip_db, st.dynamic_range,
                                    st.wf_min_db, st.wf_max_db], np.float32)


def play_buffer(x, hist, h, volume=100, balance=0.0, ratio=4):
    x = np.ascontiguousarray(x, dtype=np.int16)
    out = np.empty((ratio * x.size, 2), np.int16)
    mono = np.empty(ratio * x.size, np.float64)
    h = np.ascontiguousarray(h, np.float64)
    lib().so_play_buffer(_p(x), x.size, float(volume), float(balance), _p(h), h.size, ratio,
                         _p(hist), _p(out), _p(mono))
    return mono, out
