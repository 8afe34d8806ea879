"""CPU: the Kiwi IQ WAV reader (supersdr_b200/wavreader.py) against fixtures produced by the reference's own
kiwi/wavreader.py (oracle/make_golden.py gen_kiwi_wav) and, when the reference tree is present, against it live."""
import os
import struct
import sys

import numpy as np
import pytest

from oracle import ref_import

GOLD = os.path.join(os.path.dirname(__file__), "golden")
WAV = os.path.join(GOLD, "kiwi_iq.wav")


def test_reader_matches_reference_golden():
    import supersdr_b200 as S
    g = np.load(os.path.join(GOLD, "kiwi_iq_wav.npz"))
    t, z = S.read_kiwi_iq_wav(WAV)
    assert t.dtype == np.float64 and z.dtype == np.complex64
    assert np.array_equal(t, g["t"]) and np.array_equal(z, g["z"])
    per = list(S.KiwiIQWavReader(WAV))
    assert len(per) == int(g["n_blocks"])
    assert [tt is None for tt, _ in per] == list(g["none_blocks"])          # no time axis while the rate settles
    assert np.array_equal(per[-1][0], g["last_t"]) and np.array_equal(per[-1][1], g["last_z"])
    r = S.KiwiIQWavReader(WAV)
    for _ in r:
        pass
    assert abs(r.get_samplerate() - 12000 * (1 + 3e-5)) < 0.05


@pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")
def test_reader_matches_reference_live(tmp_path):
    import supersdr_b200 as S
    from oracle.make_golden import write_kiwi_iq_wav
    sys.path.insert(0, ref_import.REFERENCE_DIR)
    from kiwi import wavreader as ref
    rng = np.random.default_rng(3)
    blocks = [rng.integers(-32768, 32768, (n, 2)).astype(np.int16) for n in (512, 512, 300, 1024, 512, 8, 512)]
    p = str(tmp_path / "x.wav")
    write_kiwi_iq_wav(p, blocks, fs=20250, t0=99.5)
    a, b = ref.read_kiwi_iq_wav(p), S.read_kiwi_iq_wav(p)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_errors_and_source_adapter(tmp_path):
    import supersdr_b200 as S
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"RIFX" + struct.pack("<L", 4) + b"WAVE")
    with pytest.raises(S.KiwiIQWavError):
        S.KiwiIQWavReader(str(bad))
    nots = tmp_path / "nogps.wav"
    nots.write_bytes(b"RIFF" + struct.pack("<L", 36) + b"WAVE" + b"fmt " + struct.pack("<L", 16)
                     + struct.pack("<HHLLHH", 1, 2, 12000, 48000, 4, 16) + b"data" + struct.pack("<L", 4) + b"\0\0\0\0")
    with pytest.raises(S.KiwiIQWavError):
        next(S.KiwiIQWavReader(str(nots)))
    src = S.WavIQSource(WAV, wf_bins=1024)
    f1 = src.read_wf_frame()
    assert f1.dtype == np.complex64 and f1.shape == (1024,)
    assert np.all(f1.real == np.rint(f1.real)) and np.abs(f1.real).max() <= 20000      # int16 counts, kiwi/client.py:449-453
    s1, flags = src.read_snd_frame()
    assert s1.shape == (512,) and flags == 0 and np.array_equal(s1, f1[:512])
    n = 1
    while src.read_wf_frame() is not None:
        n += 1
    assert n == 3                                                                      # 6 blocks x 512 samples
