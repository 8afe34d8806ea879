"""Kiwi IQ WAV files as an offline IQ source.

Same surface and numerics as the reference's vendored reader (kiwi/wavreader.py:12-113): ``KiwiIQWavReader`` is an
iterator over ``(t, z)`` blocks -- ``z`` complex64 scaled by 1/65535, ``t`` the GNSS time of every sample from the
``kiwi`` chunk that precedes each ``data`` chunk (``None`` for the first two blocks, while the sample rate settles) --
and ``read_kiwi_iq_wav`` concatenates a file.  The RIFF walk is written directly on ``struct`` (the ``chunk`` module
the reference uses is removed in Python 3.13).  ``WavIQSource`` adapts a file to the ``iq_source`` protocol of the
drop-in classes (frames of int16-count complex64, the unit of kiwi/client.py:449-453).
"""
import collections.abc
import struct

import numpy as np


class KiwiIQWavError(Exception):
    pass


class KiwiIQWavReader(collections.abc.Iterator):
    def __init__(self, f):
        self._frame_counter = 0
        self._last_gpssec = -1
        self._f = None
        try:
            self._f = open(f, "rb")
            self._initfp()
        except Exception:
            if self._f:
                self._f.close()
            raise

    def __del__(self):
        if getattr(self, "_f", None):
            self._f.close()

    # -- RIFF plumbing ---------------------------------------------------------------------------------
    def _chunk_header(self):
        """(name, size) of the next chunk, or None at end of file (kiwi/wavreader.py:64: EOFError -> StopIteration)."""
        hdr = self._f.read(8)
        if len(hdr) < 8:
            return None
        return hdr[:4], struct.unpack("<L", hdr[4:])[0]

    def _chunk_body(self, size):
        data = self._f.read(size)
        if size & 1:
            self._f.read(1)                          # chunks are word aligned
        return data

    def _initfp(self):
        h = self._chunk_header()
        if h is None or h[0] != b"RIFF":
            raise KiwiIQWavError("file does not start with RIFF id")
        if self._f.read(4) != b"WAVE":
            raise KiwiIQWavError("not a WAVE file")
        h = self._chunk_header()
        if h is None or h[0] != b"fmt ":
            raise KiwiIQWavError("fmt chunk is missing")
        self._proc_chunk_fmt(self._chunk_body(h[1]))

    def _proc_chunk_fmt(self, body):
        # kiwi/wavreader.py:76-78
        wFormatTag, nchannels, self._samplerate, dwAvgBytesPerSec, wBlockAlign = struct.unpack("<HHLLH", body[:len(body) - 2])
        assert wFormatTag == 1 and nchannels == 2 and wBlockAlign == 4, "this is not a KiwiSDR IQ wav file"

    # -- iteration ---------------------------------------------------------------------------------------
    def __next__(self):
        return self.next()

    def next(self):
        h = self._chunk_header()
        if h is None:
            raise StopIteration
        if h[0] != b"kiwi":
            raise KiwiIQWavError("missing KiwiSDR GNSS time stamp")
        body = self._chunk_body(h[1])
        if len(body) < 10:
            raise StopIteration
        # kiwi/wavreader.py:80-82
        self.last_gps_solution, _dummy, gpssec, gpsnsec = struct.unpack("<BBII", body[:10])
        self.gpssec = gpssec + 1e-9 * gpsnsec
        h = self._chunk_header()
        if h is None:
            raise StopIteration
        if h[0] != b"data":
            raise KiwiIQWavError("missing WAVE data chunk")
        return self._proc_chunk_data(self._chunk_body(h[1]))

    def process_iq_samples(self, t, z):
        pass

    def get_samplerate(self):
        return self._samplerate

    def _proc_chunk_data(self, body):
        # kiwi/wavreader.py:84-103
        t = None
        self.last_counts = np.frombuffer(body[:len(body) & ~3], dtype=np.int16).astype(np.float32).view(np.complex64)
        z = self.last_counts / 65535
        n = len(z)
        if self._last_gpssec >= 0:
            if self._frame_counter < 3:
                self._samplerate = n / (self.gpssec - self._last_gpssec)
            else:
                self._samplerate = 0.9 * self._samplerate + 0.1 * n / (self.gpssec - self._last_gpssec)
        if self._frame_counter >= 2:
            t = np.arange(start=self.gpssec, stop=self.gpssec + (n - 0.5) / self._samplerate,
                          step=1 / self._samplerate, dtype=np.float64)
            self.process_iq_samples(t, z)
        self._last_gpssec = self.gpssec
        self._frame_counter += (self._frame_counter < 3)
        return t, z


def read_kiwi_iq_wav(filename):
    t, z = [], []
    for _t, _z in KiwiIQWavReader(filename):
        if _t is None:
            continue
        t.append(_t)
        z.append(_z)
    return np.concatenate(t), np.concatenate(z)


class WavIQSource:
    """``iq_source`` for ``kiwi_waterfall`` / ``kiwi_sound`` fed from a Kiwi IQ WAV file: the whole recording in
    int16-count units, handed out as waterfall frames of ``wf_bins`` samples and audio frames of 512 samples."""

    def __init__(self, filename, wf_bins=1024, snd_frame=512):
        r = KiwiIQWavReader(filename)
        blocks = [r.last_counts for _tz in r]                 # unscaled int16 counts of every data chunk
        self.iq = np.concatenate(blocks) if blocks else np.zeros(0, np.complex64)
        self.wf_bins, self.snd_frame = int(wf_bins), int(snd_frame)
        self._wf_pos = self._snd_pos = 0

    def read_wf_frame(self):
        if self._wf_pos + self.wf_bins > self.iq.size:
            return None
        f = self.iq[self._wf_pos:self._wf_pos + self.wf_bins]
        self._wf_pos += self.wf_bins
        return f

    def read_snd_frame(self):
        if self._snd_pos + self.snd_frame > self.iq.size:
            return None
        f = self.iq[self._snd_pos:self._snd_pos + self.snd_frame]
        self._snd_pos += self.snd_frame
        return f, 0

    def keepalive(self):
        pass

    def close(self):
        pass
