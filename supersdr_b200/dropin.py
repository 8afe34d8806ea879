"""Drop-in replacement of SuperSDR's hot path BY DELEGATION.

``bind(utils_supersdr)`` returns subclasses of the reference's own ``kiwi_waterfall`` / ``kiwi_sound`` classes: the
whole control plane -- constructor, sockets, ``SET`` messages, zoom / span / tick arithmetic, pass-band and AGC
bookkeeping, keep-alives (utils_supersdr.py:592-777,815-873,901-1029,1078-1104) -- stays the reference's code,
inherited unchanged; ONLY the hot-path methods are overridden and run on the B200 through the C ABI:

    kiwi_waterfall.receive_spectrum   utils_supersdr.py:780-785   W/F frame -> uint8 line            (a1)
    kiwi_waterfall.spectrum_db2col    utils_supersdr.py:787-813   mean + dB cal + percentiles + row  (a2, a3)
    kiwi_waterfall.run                utils_supersdr.py:879-898   averaging loop, delay deque, scroll (a2, a4)
    kiwi_waterfall.set_white_flag     utils_supersdr.py:875-877
    kiwi_waterfall.wf_data            utils_supersdr.py:692       the scrolling image (a GPU ring)    (a4)
    kiwi_sound.process_audio_stream   utils_supersdr.py:1044-1076 SND frame -> int16 PCM             (a6)
    kiwi_sound.play_buffer            utils_supersdr.py:1106-1148 x4 interpolation, balance, int16   (a9)
    kiwi_sound.set_agc_params / set_mode_freq_pb  :1022-1029      (the reference's SET message + kernel parameters)

Two data paths, chosen per frame by what arrives on the reference's own stream objects:
  * finished Kiwi frames (what a stock KiwiSDR sends): uint8 W/F lines (utils_supersdr.py:782-784) go through
    ``ssdr_wf_colorrow_u8`` (Tier P: bit-exact with the reference's arithmetic); int16 PCM SND frames
    (utils_supersdr.py:1065-1072) go straight to the GPU interpolator;
  * raw IQ (``SET mod=iq`` SND frames, kiwi/client.py:443-454): the demodulator (mix / FIR / detect / AGC) and, when an
    ``iq_source`` is attached to the waterfall object, the FFT run on the GPU too.
Nothing here imports the reference: the caller passes its module (``import utils_supersdr``) to ``bind``.
"""
import struct

import numpy as np

from .waterfall import WaterfallBank, WaterfallImage
from .sound import DemodBank, InterpBank, ResampleLine, demod_params

WF_HEADER_BYTES = 16          # "W/F" + 1 + <III (x_bin, flags|zoom, seq), utils_supersdr.py:782-783, kiwi/client.py:367-368
SND_HEADER_BYTES = 10         # "SND" + <B flags + <I seq + >H smeter, utils_supersdr.py:1065-1069
IQ_GPS_BYTES = 10             # <BBII last_gps_solution, dummy, gpssec, gpsnsec, kiwi/client.py:443-444


def parse_wf_frame(msg):
    """One W/F websocket message -> (uint8 line, x_bin, flags_zoom, seq) or None (utils_supersdr.py:782-784,
    kiwi/client.py:367-368,470-472)."""
    if not msg or bytes(msg[0:3]) != b"W/F":
        return None
    x_bin, flags_zoom, seq = struct.unpack("<III", bytes(msg[4:16]))
    return np.frombuffer(bytes(msg[WF_HEADER_BYTES:]), dtype=np.uint8), x_bin, flags_zoom, seq


def parse_snd_frame(msg):
    """One SND websocket message -> (flags, seq, rssi_dbm, payload bytes) or None (utils_supersdr.py:1065-1070)."""
    if not msg or bytes(msg[0:3]) != b"SND":
        return None
    flags, seq = struct.unpack("<BI", bytes(msg[3:8]))
    (s_meter,) = struct.unpack(">H", bytes(msg[8:10]))
    return flags, seq, 0.1 * s_meter - 127, bytes(msg[SND_HEADER_BYTES:])


class WaterfallHotPath:
    """Mixin over ``utils_supersdr.kiwi_waterfall``: the per-line work on the GPU."""

    iq_source = None              # optional: object with read_wf_frame() -> complex64[WF_BINS] (FFT path)
    _gpu_bank = None
    _gpu_bank_key = None
    _gpu_image = None
    _gpu_lines = None
    _gpu_frames = None
    _wf_cache = None

    # ---- device objects, created on first use (the reference constructor runs unchanged) --------------------
    def _bank(self, n):
        key = (int(self.WF_BINS), int(n))
        if self._gpu_bank is None or self._gpu_bank_key != key:
            if self._gpu_bank is not None:
                self._gpu_bank.close()
            self._gpu_bank = WaterfallBank(key[0], 1, key[1])
            self._gpu_bank_key = key
        return self._gpu_bank

    def _image(self):
        h = getattr(self, "_wf_height", None)
        if self._gpu_image is None or (self._gpu_image.height, self._gpu_image.width) != (h, int(self.WF_BINS)):
            if self._gpu_image is not None:
                self._gpu_image.close()
            self._gpu_image = WaterfallImage(1, h, int(self.WF_BINS))
        return self._gpu_image

    # ---- kiwi_waterfall.wf_data (utils_supersdr.py:692; read by supersdr.py:929 and plot_spectrum :1678) ----------
    @property
    def wf_data(self):
        if self._wf_cache is None:
            _, data = self._image().image(want_rgb=False, want_data=True)
            self._wf_cache = data[0]
        return self._wf_cache

    @wf_data.setter
    def wf_data(self, value):          # the reference constructor assigns np.zeros((WF_HEIGHT, WF_BINS)): (re)start the ring
        value = np.asarray(value)
        self._wf_height = int(value.shape[0])
        if self._gpu_image is not None:
            self._gpu_image.close()
            self._gpu_image = None
        self._wf_cache = np.array(value, dtype=np.float64)

    # ---- a1: utils_supersdr.py:780-785 ------------------------------------------------------------------------
    def receive_spectrum(self):
        if self._gpu_lines is None:
            self._gpu_lines, self._gpu_frames = [], []
        if self.iq_source is not None:                       # raw IQ: one frame of WF_BINS samples per line
            frame = self.iq_source.read_wf_frame()
            if frame is None:
                self.terminate = True
                return
            self._gpu_frames.append(np.asarray(frame, dtype=np.complex64).reshape(int(self.WF_BINS)))
            self.keepalive()
            return
        msg = self.wf_stream.receive_message()
        parsed = parse_wf_frame(msg)
        if parsed is not None:
            line = parsed[0]
            self._gpu_lines.append(line)
            self.spectrum = line.astype(np.float32)          # the attribute the reference sets (Kiwi byte units)
            self.kiwi_wf_seq = parsed[3]
            self.keepalive()

    # ---- a2 + a3: utils_supersdr.py:881-888 (mean) and :787-813, one fused launch ----------------------------------
    def spectrum_db2col(self):
        lines, frames = self._gpu_lines or [], self._gpu_frames or []
        self._gpu_lines, self._gpu_frames = [], []
        if frames:
            bank = self._bank(len(frames))
        elif lines:
            bank = self._bank(len(lines))
        else:                                                 # nothing new arrived: recolour the last line (GUI level changes)
            last = np.asarray(self.spectrum, dtype=np.float32)
            if not np.array_equal(last, np.rint(last)):
                return                                        # an averaged line is not a byte line: keep the previous row
            lines = [last.astype(np.uint8)]
            bank = self._bank(1)
        bank.set_display(0, 1, zoom=self.zoom, auto_scale=self.wf_auto_scaling, delta_low_db=self.delta_low_db,
                         delta_high_db=self.delta_high_db, low_clip_db=float(self.low_clip_db),
                         dynamic_range=float(self.dynamic_range))
        if frames:
            res = bank.process(np.stack(frames)[None, :, :])
        else:
            res = bank.colorrow(np.stack(lines)[None, :, :])
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

    # ---- utils_supersdr.py:875-877 -------------------------------------------------------------------------------
    def set_white_flag(self):
        self.wf_color = np.ones_like(self.wf_color) * 255
        self._image().set_white_flag()
        self._wf_cache = None

    # ---- utils_supersdr.py:879-898: one displayed line -------------------------------------------------------------
    def run_once(self):
        n = self.averaging_n if self.averaging_n > 1 else 1
        for _ in range(n):                                    # a non-W/F message in between just yields a shorter average
            self.receive_spectrum()
            if self.terminate:
                return False
        self.run_index += 1
        self.spectrum_db2col()
        if self.wf_color is None:
            return True
        # delay deque + one-line scroll (utils_supersdr.py:893-897) on the GPU ring: no O(H W) copy per line
        self._image().push(np.asarray(self.wf_color, dtype=np.float32)[None, :])
        self._wf_cache = None
        return True

    def run(self):
        while not self.terminate:
            if not self.run_once():
                break
        return


class SoundHotPath:
    """Mixin over ``utils_supersdr.kiwi_sound``: SND frames in, sound-card buffers out."""

    _gpu_demod = None
    _gpu_interp = None
    _gpu_resampler = None
    iq_mode = False               # True: the SND stream carries IQ (SET mod=iq) and the demodulator runs here

    def _demod(self):
        if self._gpu_demod is None:
            self._gpu_demod = DemodBank(1, int(self.KIWI_SAMPLES_PER_FRAME) * 4)
            self._push_demod_params()
        return self._gpu_demod

    def _push_demod_params(self):
        if self._gpu_demod is not None:
            self._gpu_demod.set_params(0, [demod_params(str(self.radio_mode), self.lc, self.hc, 0.0, self.on, self.hang, self.thresh,
                                                        self.slope, self.decay, self.gain)])

    # the reference sends the SET message; the same parameters go to the kernel (utils_supersdr.py:1022-1029)
    def set_agc_params(self):
        super().set_agc_params()
        self._push_demod_params()

    def set_mode_freq_pb(self):
        super().set_mode_freq_pb()
        self._push_demod_params()

    # ---- a6: utils_supersdr.py:1044-1076 ---------------------------------------------------------------------------
    def process_audio_stream(self):
        data = self.stream.receive_message()
        if self.run_index * self.delta_t * self.KIWI_SAMPLES_PER_FRAME / self.KIWI_RATE >= self.KIWI_SAMPLES_PER_FRAME:
            data = self.stream.receive_message()              # fractional sample-rate compensation, utils_supersdr.py:1049-1052
            self.run_index = 0
        if data is None:
            self.terminate = True
            if self.kiwi_wf is not None:
                self.kiwi_wf.terminate = True
            raise EOFError("server closed the connection")
        parsed = parse_snd_frame(data)
        if parsed is None:
            return None
        flags, _seq, rssi, payload = parsed
        self.adc_overflow_flag = True if (flags & 2) else False
        self.rssi = rssi
        if not self.iq_mode:
            return np.frombuffer(payload, dtype=">i2").astype(np.int16)
        # IQ frame (kiwi/client.py:443-454): GPS header, then big-endian int16 I,Q pairs -> K6 unpack fused into the
        # demodulator's loads (SSDR_IQ_S16BE)
        wire = np.frombuffer(payload[IQ_GPS_BYTES:], dtype=np.uint8)
        n = wire.size // 4
        res = self._demod().process(wire[:4 * n].reshape(1, n, 4), want_f32=False)
        self.rssi = float(res["rssi"][0, -1])                 # s-meter of what was demodulated here
        return res["pcm_i16"][0]

    # ---- a9: utils_supersdr.py:1106-1148 ---------------------------------------------------------------------------
    def play_buffer(self, outdata, frame_count, time_info, status):
        self.status = status
        if self.late_flag:                                    # silence right after a buffer underrun, utils_supersdr.py:1110-1115
            try:
                outdata[:] = 0
            except Exception:
                pass
            return
        popped = []
        for _ in range(self.CHUNKS):
            popped.append(self.audio_buffer.get())
        popped = np.array(popped).flatten().astype(np.int16)
        self.audio_rec_last = popped
        if self.SAMPLE_RATIO % 1:                             # high bandwidth kiwis (3ch 20kHz), utils_supersdr.py:1125-1126
            if self._gpu_resampler is None or (self._gpu_resampler.up, self._gpu_resampler.down) != (self.n_high, self.n_low):
                self._gpu_resampler = ResampleLine(self.n_high, self.n_low)
            out, mono = self._gpu_resampler.process(popped.reshape(1, -1), self.volume, self.audio_balance, want_mono=True)
        else:
            if self._gpu_interp is None:
                self._gpu_interp = InterpBank(1, int(self.SAMPLE_RATIO), taps=self.kiwi_filter.h,
                                              max_samples=int(self.KIWI_SAMPLES_PER_FRAME * self.CHUNKS))
            out, mono = self._gpu_interp.process(popped.reshape(1, -1), self.volume, self.audio_balance, want_mono=True)
        outdata[:, 0] = out[0][:, 0]
        outdata[:, 1] = out[0][:, 1]
        rec = getattr(self, "audio_rec", None)                # utils_supersdr.py:1138-1139: the recorder takes the mono buffer
        if rec is not None and getattr(rec, "recording_flag", False):
            rec.audio_buffer.append(mono[0].astype(np.int16))
        if self.rssi > self.max_rssi_before_mute:             # mute on TX, utils_supersdr.py:1141-1147
            self.mute_counter = self.muting_delay
        elif self.mute_counter > 0:
            self.mute_counter -= 1
        if self.mute_counter > 0:
            outdata *= 0
        self.old_outdata = outdata[:]


def bind(ref):
    """``ref``: the reference module (``import utils_supersdr``).  Returns a namespace with ``kiwi_waterfall`` and
    ``kiwi_sound`` subclasses whose control plane IS the reference's and whose hot path runs on the GPU, plus everything
    else of ``ref`` unchanged -- so ``from utils_supersdr import *`` in supersdr.py:6 becomes
    ``globals().update(vars(supersdr_b200.bind(utils_supersdr)))``."""
    class kiwi_waterfall(WaterfallHotPath, ref.kiwi_waterfall):
        pass

    class kiwi_sound(SoundHotPath, ref.kiwi_sound):
        pass

    kiwi_waterfall.__qualname__ = kiwi_waterfall.__name__ = "kiwi_waterfall"
    kiwi_sound.__qualname__ = kiwi_sound.__name__ = "kiwi_sound"
    ns = dict((k, v) for k, v in vars(ref).items() if not k.startswith("__"))
    ns["kiwi_waterfall"], ns["kiwi_sound"] = kiwi_waterfall, kiwi_sound
    import types
    return types.SimpleNamespace(**ns)


HOT_PATH_OVERRIDES = {
    "kiwi_waterfall": ("receive_spectrum", "spectrum_db2col", "run", "run_once", "set_white_flag", "wf_data"),
    "kiwi_sound": ("process_audio_stream", "play_buffer", "set_agc_params", "set_mode_freq_pb"),
}
