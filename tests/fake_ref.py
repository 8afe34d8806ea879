"""TEST INFRASTRUCTURE: a stand-in for the reference module ``utils_supersdr`` on hosts where the reference tree is not
present (the GPU box).  It carries ONLY what ``supersdr_b200.bind`` needs from the base classes to run the hot path:
constructors that take ready stream objects instead of opening sockets, the attributes the hot-path methods read
(names as in utils_supersdr.py:592-620,901-945), ``keepalive`` and the two ``SET`` senders.  The real control plane is
the reference's own code and is exercised by tests/test_dropin_delegation.py against the unmodified module."""
import queue
from collections import deque

import numpy as np


class filtering:
    def __init__(self, fl, fs):
        b = fl / fs
        N = int(np.ceil(4 / b))
        if not N % 2:
            N += 1
        k = np.arange(N)
        h = np.sinc(2 * fl / fs * (k - (N - 1) / 2)) * np.blackman(N)
        self.h, self.n_tap = h / np.sum(h), N


class kiwi_waterfall:
    MAX_ZOOM, WF_BINS, MIN_DYN_RANGE = 14, 1024, 40.
    CLIP_LOWP, CLIP_HIGHP = 40., 100
    delta_low_db, delta_high_db = 0, 0
    low_clip_db, high_clip_db = -120, -60
    wf_min_db, wf_max_db = -120, -80
    wf_buffer_len = 3

    def __init__(self, wf_stream, zoom_, disp, wf_bins=1024):
        self.wf_stream = wf_stream
        self.zoom = zoom_
        self.WF_BINS = wf_bins
        self.averaging_n = 1
        self.wf_auto_scaling = True
        self.dynamic_range = self.MIN_DYN_RANGE
        self.terminate = False
        self.run_index = 0
        self.wf_color = None
        self.keepalives = 0
        self.wf_data = np.zeros((disp.WF_HEIGHT, self.WF_BINS))
        self.wf_data_tmp = deque([], self.wf_buffer_len)

    def keepalive(self):
        self.keepalives += 1
        self.wf_stream.send_message("SET keepalive")


class kiwi_sound:
    FORMAT, CHANNELS, AUDIO_RATE, KIWI_RATE = np.int16, 2, 48000, 12000
    SAMPLE_RATIO = int(AUDIO_RATE / KIWI_RATE)
    CHUNKS, KIWI_SAMPLES_PER_FRAME = 1, 512

    def __init__(self, stream, mode_, lc_, hc_, kiwi_wf, buffer_len, volume_=100):
        self.stream, self.kiwi_wf = stream, kiwi_wf
        self.FULL_BUFF_LEN = max(1, buffer_len)
        self.audio_buffer = queue.Queue(maxsize=self.FULL_BUFF_LEN)
        self.terminate, self.volume = False, volume_
        self.max_rssi_before_mute, self.mute_counter, self.muting_delay = -20, 0, 15
        self.adc_overflow_flag, self.status, self.run_index, self.delta_t, self.rssi = False, None, 0, 0.0, -127
        self.freq, self.radio_mode, self.lc, self.hc = 7100, mode_, lc_, hc_
        self.on, self.hang, self.thresh, self.slope, self.decay, self.gain = True, False, -80, 0, 4000, 50
        self.audio_balance, self.late_flag = 0.0, False
        self.kiwi_filter = filtering(self.KIWI_RATE / 2, self.AUDIO_RATE)
        self.n_tap = self.kiwi_filter.n_tap
        gcd = np.gcd(self.KIWI_RATE, self.AUDIO_RATE)
        self.n_low, self.n_high = int(self.KIWI_RATE / gcd), int(self.AUDIO_RATE / gcd)
        self.audio_rec = type("rec", (), {"recording_flag": False, "audio_buffer": []})()

    def set_agc_params(self):
        self.stream.send_message("SET agc=%d hang=%d thresh=%d slope=%d decay=%d manGain=%d" % (
            self.on, self.hang, self.thresh, self.slope, self.decay, self.gain))

    def set_mode_freq_pb(self):
        self.stream.send_message("SET mod=%s low_cut=%d high_cut=%d freq=%.3f" % (self.radio_mode.lower(), self.lc, self.hc, self.freq))

    def keepalive(self):
        self.stream.send_message("SET keepalive")
