"""CPU: the drop-ins are made BY DELEGATION: bound to the UNMODIFIED reference module, the classes
inherit its whole control plane and override only the hot-path methods; the product's wire-frame parsing gives what the
reference's own receive functions give on the same bytes.  Needs /root/reference (build container only)."""
import struct
import types

import numpy as np
import pytest

import fake_kiwi
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def _bound():
    import supersdr_b200 as S
    ref = ref_import.load()
    return S, ref, S.bind(ref)


def test_control_plane_is_the_references_own_code():
    S, ref, K = _bound()
    for cls_name, overrides in S.HOT_PATH_OVERRIDES.items():
        cls, base = getattr(K, cls_name), getattr(ref, cls_name)
        assert issubclass(cls, base) and cls.__name__ == cls_name
        own = set()
        for klass in cls.__mro__:
            if klass is base:
                break
            own |= {k for k in vars(klass) if not k.startswith("_")}
        own -= {"iq_source", "iq_mode"}
        assert own == set(overrides), (cls_name, own ^ set(overrides))
        # everything else resolves to the reference's function objects
        for name, fn in vars(base).items():
            if callable(fn) and not name.startswith("__") and name not in overrides:
                assert getattr(cls, name) is fn, name
    for name in ("set_freq_zoom", "zoom_to_span", "start_frequency_to_counter", "bins_to_khz", "gen_div", "change_passband", "keepalive"):
        assert getattr(K.kiwi_waterfall, name) is getattr(ref.kiwi_waterfall, name)
    for name in ("change_passband", "change_agc_delay", "run", "get_audio_chunk", "keepalive", "close_connection"):
        assert getattr(K.kiwi_sound, name) is getattr(ref.kiwi_sound, name)
    # the rest of the module passes through: `from utils_supersdr import *` keeps working
    assert K.filtering is ref.filtering and K.display_stuff is ref.display_stuff


def test_wire_parsing_equals_the_references():
    """parse_wf_frame / parse_snd_frame against kiwi_waterfall.receive_spectrum (utils_supersdr.py:780-785) and
    kiwi_sound.process_audio_stream (:1044-1076) of the unmodified reference on the same bytes."""
    S, ref, K = _bound()
    rng = np.random.default_rng(0)
    line = rng.integers(0, 256, 1024).astype(np.uint8)
    msg = fake_kiwi.wf_frame(line, x_bin=123, zoom=5, seq=77)
    assert len(msg) == 16 + 1024
    r = ref.kiwi_waterfall.__new__(ref.kiwi_waterfall)
    r.wf_stream = fake_kiwi.FakeKiwiStream([msg])
    r.keepalive = lambda: None
    r.receive_spectrum()
    got = S.parse_wf_frame(msg)
    assert np.array_equal(got[0].astype(np.float32), r.spectrum) and got[1:] == (123, 5, 77)
    assert S.parse_wf_frame(b"MSG x=1") is None and S.parse_wf_frame(None) is None
    # the bound class's receive_spectrum sets the same attribute from the same bytes (host side only, no GPU)
    k = K.kiwi_waterfall.__new__(K.kiwi_waterfall)
    k.wf_stream = fake_kiwi.FakeKiwiStream([msg])
    k.keepalive = lambda: None
    k.receive_spectrum()
    assert np.array_equal(k.spectrum, r.spectrum) and k.spectrum.dtype == np.float32

    pcm = rng.integers(-32768, 32768, 512).astype(np.int16)
    smsg = fake_kiwi.snd_frame(pcm, rssi_dbm=-61.3, flags=2, seq=9)
    assert len(smsg) == 10 + 1024
    rs = ref.kiwi_sound.__new__(ref.kiwi_sound)
    rs.stream = fake_kiwi.FakeKiwiStream([smsg])
    rs.run_index, rs.delta_t, rs.KIWI_SAMPLES_PER_FRAME, rs.KIWI_RATE = 0, 0.0, 512, 12000
    ref_pcm = rs.process_audio_stream()
    flags, seq, rssi, payload = S.parse_snd_frame(smsg)
    assert np.array_equal(np.frombuffer(payload, ">i2").astype(np.int16), ref_pcm)
    assert abs(rssi - rs.rssi) < 1e-12 and bool(flags & 2) == rs.adc_overflow_flag and seq == 9
    ks = K.kiwi_sound.__new__(K.kiwi_sound)
    ks.stream = fake_kiwi.FakeKiwiStream([smsg])
    ks.run_index, ks.delta_t, ks.KIWI_SAMPLES_PER_FRAME, ks.KIWI_RATE, ks.kiwi_wf = 0, 0.0, 512, 12000, None
    assert np.array_equal(ks.process_audio_stream(), ref_pcm) and ks.rssi == rs.rssi and ks.adc_overflow_flag


def test_iq_frame_layout_matches_kiwiclient():
    """fake_kiwi.iq_frame lays the bytes out as kiwi/client.py:443-454 reads them."""
    iq = (np.arange(16) * 100 - 700) + 1j * (np.arange(16) * -50 + 3)
    msg = fake_kiwi.iq_frame(iq.astype(np.complex64), gpssec=5, gpsnsec=6)
    data = msg[10:]
    gps = struct.unpack("<BBII", data[0:10])
    assert gps == (0, 0, 5, 6)
    samples = np.ndarray(len(data[10:]) // 2, dtype=">h", buffer=data[10:]).astype(np.float32)
    cs = samples[0::2] + 1j * samples[1::2]
    assert np.array_equal(cs, iq)
